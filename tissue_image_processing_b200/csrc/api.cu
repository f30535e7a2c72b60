// C ABI of libtsp_b200 (see include/tsp_b200.h for the reference interfaces each entry replaces).
#include <sched.h>
#include <stdarg.h>

#include <algorithm>
#include <condition_variable>
#include <thread>

#include "common.cuh"

namespace tsp {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
}

thread_local bool tl_chain_launches = true;
std::atomic<bool> g_no_chain{false};
std::atomic<int> g_chain_plain_mask{0};
thread_local int tl_chain_site = 0;
bool g_sync_launches = false;

static const char* kStageNames[STG_COUNT] = {"percentile", "decimate", "coarse", "interp_argmax", "prepare",
                                             "blur_pre", "blur_score", "argmax", "band", "widen",
                                             "percentile_sample", "percentile_count"};

void prof_mark(tsp_handle* h, cudaStream_t s, int stage) {
    if (!h->profiling) return;
    cudaEvent_t ev;
    if (!h->prof_pool.empty()) {
        ev = h->prof_pool.back();
        h->prof_pool.pop_back();
    } else if (cudaEventCreate(&ev) != cudaSuccess) {
        return;
    }
    cudaEventRecord(ev, s);
    h->prof_events.push_back(ev);
    h->prof_stage.push_back(stage);
}

int get_tensor_map_encoder(EncodeTiledFn* out) {
    static EncodeTiledFn encode = nullptr;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        TSP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !fn) {
            set_error("cuTensorMapEncodeTiled is not available from this driver");
            return TSP_ERR_CUDA;
        }
        encode = (EncodeTiledFn)fn;
    }
    *out = encode;
    return TSP_OK;
}

struct Crop {
    int z0;        // first plane of the cropped stack inside the full stack
    int zc;        // planes after SP:30-31
    int z_offset;  // added to the argmax (SP:61 adds min_z even when max_z == 0)
};

static int resolve_crop(const tsp_frame_desc* d, Crop* c) {
    if (d->channels < 1 || d->planes < 1 || d->rows < 1 || d->cols < 1) {
        set_error("invalid shape C=%d Z=%d Y=%d X=%d", d->channels, d->planes, d->rows, d->cols);
        return TSP_ERR_INVALID;
    }
    if (d->reference_channel < 0 || d->reference_channel >= d->channels) {
        set_error("reference_channel %d out of range for %d channels", d->reference_channel, d->channels);
        return TSP_ERR_INVALID;
    }
    if (d->min_z < 0 || d->max_z < 0) {
        set_error("negative min_z/max_z are not supported");
        return TSP_ERR_INVALID;
    }
    if (d->mode < TSP_MODE_FAST || d->mode > TSP_MODE_BITEXACT) {
        set_error("unknown mode %d", d->mode);
        return TSP_ERR_INVALID;
    }
    if (d->bin_size > 1 && (d->method < TSP_METHOD_MAX_AVERAGES || d->method > TSP_METHOD_MULTI_CHANNEL)) {
        set_error("unknown method %d", d->method);
        return TSP_ERR_INVALID;
    }
    if ((d->flags & ~(TSP_FRAME_CONCURRENT | TSP_FRAME_OUT_U16)) != 0 || d->reserved != 0 ||
        (d->has_params != 0 && d->has_params != 1)) {
        set_error("unknown flags 0x%x (or non-zero reserved words) in the frame descriptor", (unsigned)d->flags);
        return TSP_ERR_INVALID;
    }
    c->z_offset = d->min_z;
    if (d->max_z > 0) {
        const int hi = d->max_z < d->planes ? d->max_z : d->planes;
        c->z0 = d->min_z;
        c->zc = hi - d->min_z;
        if (c->zc < 1) {
            set_error("empty z crop [%d:%d) of %d planes", d->min_z, d->max_z, d->planes);
            return TSP_ERR_INVALID;
        }
    } else {
        c->z0 = 0;
        c->zc = d->planes;
    }
    return TSP_OK;
}

static const tsp_params kReferenceParams = {95.0f, kAiryscanPedestal, {0.5f, 1.0f, 1.0f}, {0.5f, 30.0f, 30.0f},
                                            {1.0f, 2.0f, 2.0f}, {0, 0, 0, 0, 0}};

// desc.params (or the reference constants) -> what the stages take, and whether the fast-mode tables / fused band
// kernels (built for the reference's sigmas) apply
static int resolve_params(const tsp_frame_desc* d, Params* out) {
    const tsp_params& p = d->has_params ? d->params : kReferenceParams;
    if (!(p.percentile >= 0.0f && p.percentile <= 100.0f)) {
        set_error("percentile %g outside [0, 100]", (double)p.percentile);        // numpy: "Percentiles must be in the range [0, 100]"
        return TSP_ERR_INVALID;
    }
    if (p.pedestal < 0 || p.pedestal > 65535) {
        set_error("pedestal %d outside [0, 65535]", p.pedestal);
        return TSP_ERR_INVALID;
    }
    for (int i = 0; i < 5; ++i)
        if (p.reserved[i] != 0) {
            set_error("non-zero reserved words in tsp_params");
            return TSP_ERR_INVALID;
        }
    out->q = p.percentile / 100.0f;            // float32 division, as numpy forms q for float32 data
    out->q64 = (double)p.percentile / 100.0;
    out->pedestal = d->airyscan ? p.pedestal : 0;
    out->default_score = true;
    out->default_mask = true;
    for (int i = 0; i < 3; ++i) {
        if (!(p.sigma_pre[i] >= 0.0f) || !(p.sigma_score[i] >= 0.0f) || !(p.sigma_mask[i] >= 0.0f) ||
            p.sigma_pre[i] > 512.0f || p.sigma_score[i] > 512.0f || p.sigma_mask[i] > 512.0f) {
            set_error("sigmas must lie in [0, 512]");
            return TSP_ERR_INVALID;
        }
        out->sigma_pre[i] = (double)p.sigma_pre[i];
        out->sigma_score[i] = (double)p.sigma_score[i];
        out->sigma_mask[i] = (double)p.sigma_mask[i];
        if (p.sigma_pre[i] != kReferenceParams.sigma_pre[i] || p.sigma_score[i] != kReferenceParams.sigma_score[i])
            out->default_score = false;
        if (p.sigma_mask[i] != kReferenceParams.sigma_mask[i]) out->default_mask = false;
    }
    return TSP_OK;
}

struct Workspace {
    int32_t* status;
    uint32_t* hist;
    int* worklist;
    float* volA;
    float* volB;
    void* fast;
    // binned scores / manifold (general path)
    float* binned;        // (zc, cy, cx) score
    int32_t* status2;     // status block of the second channel's percentile (SP:46)
    int32_t* coarse_z;    // (cy, cx) height map grown on the binned score
    int32_t* zmap_other;  // (Y, X) height map of the other channels when it is resampled separately
    void* mf_scratch;
    size_t total;
};

static bool general_path(const tsp_frame_desc* d) { return d->bin_size > 1 || d->build_manifold != 0; }
static bool fast_score(const tsp_frame_desc* d, const Params& pr) {
    return d->mode == TSP_MODE_FAST && !general_path(d) && pr.default_score;
}
// the fused band kernels carry the reference's (1, 2, 2) mask and fp32 arithmetic; everything else materialises it
static bool fused_band(const tsp_frame_desc* d, const Params& pr) {
    return d->mode != TSP_MODE_BITEXACT && pr.default_mask;
}

static Workspace carve(const tsp_frame_desc* d, const Crop& c, const Params& pr, void* base) {
    Workspace w{};
    char* p = (char*)base;
    size_t off = 0;
    w.status = (int32_t*)(p + off);
    off += align_up(kStatusWords * sizeof(int32_t), 256);
    w.hist = (uint32_t*)(p + off);
    off += align_up(percentile_scratch_bytes(), 256);
    w.worklist = (int*)(p + off);
    off += align_up(band_worklist_bytes(d->rows, d->cols), 256);
    const size_t vol = align_up((size_t)c.zc * d->rows * d->cols * sizeof(float), 256);
    if (general_path(d)) {
        w.volA = (float*)(p + off);
        off += vol;
        w.volB = (float*)(p + off);
        off += vol;
        const int bin = d->bin_size > 1 ? d->bin_size : 1;
        const size_t cy = (d->rows + bin - 1) / bin, cx = (d->cols + bin - 1) / bin;
        w.binned = (float*)(p + off);
        off += align_up((size_t)c.zc * cy * cx * sizeof(float), 256);
        w.status2 = (int32_t*)(p + off);
        off += align_up(kStatusWords * sizeof(int32_t), 256);
        w.coarse_z = (int32_t*)(p + off);
        off += align_up(cy * cx * sizeof(int32_t), 256);
        w.zmap_other = (int32_t*)(p + off);
        off += align_up((size_t)d->rows * d->cols * sizeof(int32_t), 256);
        w.mf_scratch = p + off;
        off += align_up(manifold_scratch_bytes(), 256);
    } else {
        if (fast_score(d, pr)) {
            w.fast = p + off;
            off += fast_workspace_bytes(c.zc, d->rows, d->cols);
        }
        if (!fast_score(d, pr) || !fused_band(d, pr)) {
            w.volA = (float*)(p + off);
            off += vol;
            w.volB = (float*)(p + off);
            off += vol;
        }
    }
    w.total = off;
    return w;
}

// ---- host buffers in pageable memory --------------------------------------------------------------
// cudaMemcpyAsync from pageable memory is staged by the driver on one thread: 11 GB/s measured on the B200 box, 47 ms
// for a 512 MiB stack the link moves in 10.  A caller of the reference's operator hands over whatever its reader
// returned (an ordinary numpy array), so the library stages such a stack itself: 32 MiB chunks go through a ring of
// pinned buffers, each chunk copied by a pool of host threads (48 GB/s on 16 threads) while the DMA of the chunk
// before it runs.  Pinned / registered buffers take the direct path.
struct CopyPool {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    char* dst = nullptr;
    const char* src = nullptr;
    size_t bytes = 0, pending = 0;
    uint64_t generation = 0;
    bool quit = false;

    explicit CopyPool(int n) {
        for (int i = 0; i < n; ++i) workers.emplace_back([this, i] { loop(i); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
        }
        cv_work.notify_all();
        for (auto& t : workers) t.join();
    }
    std::mutex call_mu;          // one copy at a time (callers of different frame slots may arrive together)
    void copy(void* d, const void* s, size_t n) {
        if (workers.size() < 2 || n < ((size_t)1 << 20)) {
            memcpy(d, s, n);
            return;
        }
        std::lock_guard<std::mutex> turn(call_mu);
        std::unique_lock<std::mutex> lk(mu);
        dst = (char*)d;
        src = (const char*)s;
        bytes = n;
        pending = workers.size();
        ++generation;
        cv_work.notify_all();
        cv_done.wait(lk, [&] { return pending == 0; });
    }
    void loop(int i) {
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu);
            cv_work.wait(lk, [&] { return quit || generation != seen; });
            if (quit) return;
            seen = generation;
            char* d = dst;
            const char* s = src;
            const size_t n = bytes, parts = workers.size();
            lk.unlock();
            const size_t per = ((n + parts - 1) / parts + 4095) & ~(size_t)4095;
            const size_t lo = std::min(n, (size_t)i * per), hi = std::min(n, lo + per);
            if (hi > lo) memcpy(d + lo, s + lo, hi - lo);
            lk.lock();
            if (--pending == 0) cv_done.notify_one();
        }
    }
};

static int env_int_or(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

static int copy_pool_threads() {
    int n = env_int_or("TSP_COPY_THREADS", 0);
    if (n <= 0) {
        cpu_set_t set;
        CPU_ZERO(&set);
        int cpus = sched_getaffinity(0, sizeof set, &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
        int share = env_int_or("LOCAL_WORLD_SIZE", 1);          // the ranks of a torchrun node share its CPUs
        if (share < 1) share = 1;
        n = cpus / share;
    }
    return n < 1 ? 1 : (n > 16 ? 16 : n);
}

static bool host_pointer_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

constexpr size_t kStageChunk = (size_t)32 << 20;

// d_dst[0, bytes) <- pageable h_src, enqueued on `s`; returns when the last chunk has been handed to the DMA engine
static int staged_copy_in(tsp_handle* h, tsp_handle::Slot& sl, char* d_dst, const char* h_src, size_t bytes,
                          cudaStream_t s) {
    if (!h->copy_pool) h->copy_pool = new CopyPool(copy_pool_threads());
    const int ring = tsp_handle::Slot::kStageRing;
    for (int b = 0; b < ring; ++b) {
        if (!sl.stage[b]) TSP_CUDA(cudaMallocHost(&sl.stage[b], kStageChunk));
        if (!sl.stage_ev[b]) TSP_CUDA(cudaEventCreateWithFlags(&sl.stage_ev[b], cudaEventDisableTiming));
    }
    size_t off = 0;
    for (int c = 0; off < bytes; ++c, off += kStageChunk) {
        const int b = c % ring;
        const size_t n = std::min(kStageChunk, bytes - off);
        if (c >= ring) TSP_CUDA(cudaEventSynchronize(sl.stage_ev[b]));        // the DMA that read this buffer is done
        h->copy_pool->copy(sl.stage[b], h_src + off, n);
        TSP_CUDA(cudaMemcpyAsync(d_dst + off, sl.stage[b], n, cudaMemcpyHostToDevice, s));
        TSP_CUDA(cudaEventRecord(sl.stage_ev[b], s));
    }
    return TSP_OK;
}

}  // namespace tsp

using namespace tsp;

extern "C" {

int tsp_abi_version(void) { return TSP_ABI_VERSION; }

const char* tsp_last_error(void) { return g_error; }

int tsp_create(int device, tsp_handle** out) {
    if (!out) return TSP_ERR_INVALID;
    int count = 0;
    TSP_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) {
        set_error("device %d not present (%d visible)", device, count);
        return TSP_ERR_INVALID;
    }
    TSP_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    TSP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("libtsp_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
        return TSP_ERR_INVALID;
    }
    tsp_handle* h = new tsp_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    TSP_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    TSP_CUDA(cudaMallocHost((void**)&h->h_status, kStatusWords * sizeof(int32_t)));
    // kernel-variant switches: the environment is read here, once (tsp_debug_set changes them later)
    auto env_int = [](const char* name, int dflt) {
        const char* v = getenv(name);
        return v && *v ? atoi(v) : dflt;
    };
    h->dbg.no_ring = env_int("TSP_NO_RING", 0);
    h->dbg.band_variant = env_int("TSP_BAND_VARIANT", 0);
    h->dbg.interp_rows = env_int("TSP_INTERP_ROWS", 4);
    h->dbg.graphs = env_int("TSP_NO_GRAPHS", 0) ? 0 : 1;
    h->dbg.interp_global = env_int("TSP_INTERP_GLOBAL", 0);
    if (env_int("TSP_NO_CHAIN", 0)) g_no_chain.store(true);
    g_sync_launches = env_int("TSP_SYNC_LAUNCHES", 0) != 0;
    *out = h;
    return TSP_OK;
}

void tsp_default_params(tsp_params* out) {
    if (out) *out = kReferenceParams;
}

int tsp_debug_set(tsp_handle* h, const char* key, int value) {
    if (!h || !key) return TSP_ERR_INVALID;
    if (!strcmp(key, "no_ring")) h->dbg.no_ring = value;
    else if (!strcmp(key, "band_variant") && (value == 0 || value == 2 || value == 3)) h->dbg.band_variant = value;
    else if (!strcmp(key, "interp_rows") && (value == 2 || value == 4 || value == 8)) h->dbg.interp_rows = value;
    else if (!strcmp(key, "no_chain")) g_no_chain.store(value != 0);
    else if (!strcmp(key, "chain_plain_mask")) g_chain_plain_mask.store(value);
    else if (!strcmp(key, "graphs")) h->dbg.graphs = value != 0;
    else if (!strcmp(key, "interp_global")) h->dbg.interp_global = value != 0;
    else {
        set_error("unknown debug switch %s = %d", key, value);
        return TSP_ERR_INVALID;
    }
    return TSP_OK;
}

int tsp_destroy(tsp_handle* h) {
    if (!h) return TSP_OK;
    cudaSetDevice(h->device);
    for (auto& kv : h->taps) {
        cudaFree((void*)kv.second.w64);
        cudaFree((void*)(kv.second.w32 - kTapPad));
    }
    for (auto& kv : h->tables) cudaFree(kv.second);
    if (h->d_scratch) cudaFree(h->d_scratch);
    for (auto& sl : h->slots) {
        if (sl.stream) cudaStreamSynchronize(sl.stream);
        if (sl.d_mem) cudaFree(sl.d_mem);
        if (sl.h_status) cudaFreeHost(sl.h_status);
        if (sl.h_out) cudaFreeHost(sl.h_out);
        for (int b = 0; b < tsp_handle::Slot::kStageRing; ++b) {
            if (sl.stage[b]) cudaFreeHost(sl.stage[b]);
            if (sl.stage_ev[b]) cudaEventDestroy(sl.stage_ev[b]);
        }
        if (sl.stream) cudaStreamDestroy(sl.stream);
    }
    delete h->copy_pool;
    for (auto& g : h->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    if (h->capture_stream) cudaStreamDestroy(h->capture_stream);
    if (h->h_status) cudaFreeHost(h->h_status);
    if (h->stream) cudaStreamDestroy(h->stream);
    for (cudaEvent_t ev : h->prof_events) cudaEventDestroy(ev);
    for (cudaEvent_t ev : h->prof_pool) cudaEventDestroy(ev);
    delete h;
    return TSP_OK;
}

int64_t tsp_launch_count(const tsp_handle* h) { return h ? h->launches : 0; }

size_t tsp_project_workspace_bytes(const tsp_frame_desc* desc) {
    Crop c;
    Params pr;
    if (!desc || resolve_crop(desc, &c) || resolve_params(desc, &pr)) return 0;
    return carve(desc, c, pr, nullptr).total;
}

}  // extern "C"

// Enqueues the kernels of one frame on `s` (plain launches or, when `s` is being captured, graph nodes).
static int project_frame_eager(tsp_handle* h, const tsp_frame_desc* desc, const uint16_t* d_stack, float* d_proj,
                               int32_t* d_zmap, void* d_workspace, size_t workspace_bytes, cudaStream_t s) {
    Crop c;
    Params pr;
    int rc = resolve_crop(desc, &c);
    if (rc) return rc;
    rc = resolve_params(desc, &pr);
    if (rc) return rc;
    Workspace w = carve(desc, c, pr, d_workspace);
    if (workspace_bytes < w.total) {
        set_error("workspace %zu bytes < required %zu", workspace_bytes, w.total);
        return TSP_ERR_WORKSPACE;
    }
    struct ChainScope {          // launches happen on this thread: the flag covers exactly this frame's kernels
        explicit ChainScope(bool on) { tl_chain_launches = on; tl_chain_site = 0; }
        ~ChainScope() { tl_chain_launches = true; }
    } chain_scope((desc->flags & TSP_FRAME_CONCURRENT) == 0);
    const int Y = desc->rows, X = desc->cols, C = desc->channels;
    const size_t plane = (size_t)Y * X;
    const size_t chan_stride = (size_t)desc->planes * plane;
    const size_t z0_off = (size_t)c.z0 * plane;
    const size_t nvox = (size_t)c.zc * plane;
    const int ped = pr.pedestal;
    const uint16_t* ref = d_stack + (size_t)desc->reference_channel * chan_stride + z0_off;
    const double* sig_pre = pr.sigma_pre;          // SP:37
    const double* sig_score = pr.sigma_score;      // SP:55
    const bool band_fused = fused_band(desc, pr);

    int* worklist = w.worklist;
    NvtxRange frame_range("tsp/frame");
    prof_mark(h, s, -1);
    if (general_path(desc)) {
        // SP:39-65 with bin_size > 1 and / or build_manifold: direct-FIR score volumes, binned, then either the
        // resampled argmax or the region growing
        const bool fp64 = desc->mode == TSP_MODE_BITEXACT;
        const int bin = desc->bin_size > 1 ? desc->bin_size : 1;
        const bool binned = bin > 1, manifold = desc->build_manifold != 0;
        const int cy = (Y + bin - 1) / bin, cx = (X + bin - 1) / bin;
        rc = launch_percentile(h, ref, nvox, ped, w.status, w.hist, s, nullptr, 0, pr.q, pr.q64);
        if (rc) return rc;
        prof_mark(h, s, STG_PERCENTILE);
        NvtxRange general_range("tsp/general_score");
        rc = launch_prepare(h, ref, w.volA, nvox, ped, w.status, s);
        if (rc) return rc;
        rc = gaussian_blur<float>(h, w.volA, w.volB, w.volA, c.zc, Y, X, sig_pre, fp64, s);     // volB = pc
        if (rc) return rc;
        const float* score = nullptr;
        if (!binned || desc->method == TSP_METHOD_MAX_AVERAGES) {
            rc = gaussian_blur<float>(h, w.volB, w.volA, w.volB, c.zc, Y, X, sig_score, fp64, s);
            if (rc) return rc;
            score = w.volA;
            if (binned) {
                rc = launch_block_reduce(h, w.volA, w.binned, c.zc, Y, X, bin, false, false, s);
                if (rc) return rc;
                score = w.binned;
            }
        } else {
            rc = launch_block_reduce(h, w.volB, w.binned, c.zc, Y, X, bin, true, false, s);    // SP:43 / SP:49
            if (rc) return rc;
            score = w.binned;
            if (desc->method == TSP_METHOD_MULTI_CHANNEL) {                                     // SP:44-51
                const int oc = (desc->reference_channel + 1) % C;
                const uint16_t* other = d_stack + (size_t)oc * chan_stride + z0_off;
                rc = launch_percentile_all(h, other, nvox, ped, w.status2, w.hist, s, pr.q);
                if (rc) return rc;
                rc = launch_prepare(h, other, w.volA, nvox, ped, w.status2, s);
                if (rc) return rc;
                rc = gaussian_blur<float>(h, w.volA, w.volB, w.volA, c.zc, Y, X, sig_pre, fp64, s);
                if (rc) return rc;
                rc = gaussian_blur<float>(h, w.volB, w.volA, w.volB, c.zc, Y, X, sig_score, fp64, s);
                if (rc) return rc;
                rc = launch_block_reduce(h, w.volA, w.binned, c.zc, Y, X, bin, false, true, s);
                if (rc) return rc;
            }
        }
        prof_mark(h, s, STG_BLUR_SCORE);
        const int32_t* zmap_other = nullptr;
        bool range_known = true;
        if (!manifold) {
            rc = launch_resize_argmax(h, score, d_zmap, c.zc, Y, X, cy, cx, c.z_offset, w.status, s);
            if (rc) return rc;
        } else if (!binned) {
            rc = launch_manifold(h, score, d_zmap, c.zc, Y, X, w.status, w.mf_scratch, s);      // SP:57: no min_z added
            if (rc) return rc;
        } else {
            rc = launch_manifold(h, score, w.coarse_z, c.zc, cy, cx, w.status, w.mf_scratch, s);
            if (rc) return rc;
            rc = launch_resize_round(h, w.coarse_z, d_zmap, Y, X, cy, cx, 0, 0, s);            // SP:64
            if (rc) return rc;
            range_known = false;
            if (desc->atoh_shift != 0 && C > 1) {
                rc = launch_resize_round(h, w.coarse_z, w.zmap_other, Y, X, cy, cx, desc->atoh_shift, c.zc, s);   // SP:62, 65
                if (rc) return rc;
                zmap_other = w.zmap_other;
            }
        }
        prof_mark(h, s, STG_ARGMAX);
        NvtxRange band_range("tsp/band");
        if (!band_fused)
            rc = launch_band_project_bitexact_ex(h, d_stack, chan_stride, z0_off, d_zmap, d_proj, C, c.zc, Y, X,
                                                 desc->reference_channel, desc->atoh_shift, ped, w.volA, w.volB,
                                                 w.status, range_known, s, zmap_other, pr.sigma_mask, fp64);
        else
            rc = launch_band_project_ex(h, d_stack, chan_stride, z0_off, d_zmap, d_proj, C, c.zc, Y, X,
                                        desc->reference_channel, desc->atoh_shift, ped, w.status, range_known, s,
                                        worklist, zmap_other);
        prof_mark(h, s, STG_BAND);
        return rc;
    }
    const bool fast = fast_score(desc, pr);
    const bool fp64 = desc->mode == TSP_MODE_BITEXACT;
    {
        NvtxRange r("tsp/percentile");
        rc = launch_percentile(h, ref, nvox, ped, w.status, w.hist, s, fast ? w.fast : nullptr,
                               fast ? fast_accum_bytes(c.zc, Y, X) : 0, pr.q, pr.q64);
        if (rc) return rc;
        prof_mark(h, s, STG_PERCENTILE);
    }
    if (fast) {
        rc = launch_fast_score_argmax(h, ref, d_zmap, c.zc, Y, X, ped, c.z_offset, w.status, w.fast, s, true);
        if (rc) return rc;
    } else {
        NvtxRange r("tsp/score_fir");
        if (fp64) {
            rc = launch_prepare(h, ref, w.volA, nvox, ped, w.status, s);
            if (rc) return rc;
            prof_mark(h, s, STG_PREPARE);
            rc = gaussian_blur<float>(h, w.volA, w.volB, w.volA, c.zc, Y, X, sig_pre, true, s);
        } else {                       // fp32: conversion, pedestal and clip ride in the first (z) pass
            rc = prepare_and_blur_f32(h, ref, w.volB, w.volA, c.zc, Y, X, sig_pre, ped, w.status, s);
        }
        if (rc) return rc;
        prof_mark(h, s, STG_BLUR_PRE);
        rc = gaussian_blur<float>(h, w.volB, w.volA, w.volB, c.zc, Y, X, sig_score, fp64, s);
        if (rc) return rc;
        prof_mark(h, s, STG_BLUR_SCORE);
        rc = launch_argmax(h, w.volA, d_zmap, c.zc, Y, X, c.z_offset, w.status, s);
        if (rc) return rc;
        prof_mark(h, s, STG_ARGMAX);
    }
    NvtxRange band_range("tsp/band");
    if (!band_fused)
        rc = launch_band_project_bitexact_ex(h, d_stack, chan_stride, z0_off, d_zmap, d_proj, C, c.zc, Y, X,
                                             desc->reference_channel, desc->atoh_shift, ped, w.volA, w.volB,
                                             w.status, true, s, nullptr, pr.sigma_mask, fp64);
    else
        rc = launch_band_project_ex(h, d_stack, chan_stride, z0_off, d_zmap, d_proj, C, c.zc, Y, X,
                                    desc->reference_channel, desc->atoh_shift, ped, w.status, true, s, worklist);
    prof_mark(h, s, STG_BAND);
    return rc;
}


// ---- CUDA graph replay ------------------------------------------------------------------------------------------
// A frame is a dozen launches (memset, tensor-map encodes, ten kernels): ~40 us of host time, more than the GPU needs
// for a 512x512x32 stack.  Frames of a movie run again and again with the same descriptor and the same buffers (a
// frame slot, a DeviceProjector), so the launch sequence is captured once into a CUDA graph - programmatic dependent
// launches become programmatic edges - and replayed with one cudaGraphLaunch.  Keyed by descriptor + pointers; the
// first call with a key runs eagerly (it also uploads per-shape tables, which cannot be captured), the second is
// captured, later ones replay.  Anything that fails to capture falls back to plain launches for good.
struct GraphKey {
    tsp_frame_desc desc;
    const void* stack;
    void* proj;
    void* zmap;
    void* ws;
    size_t ws_bytes;
};

static std::string graph_key(const tsp_frame_desc* desc, const void* stack, void* proj, void* zmap, void* ws, size_t n) {
    GraphKey k;
    memset(&k, 0, sizeof k);
    k.desc = *desc;
    k.stack = stack; k.proj = proj; k.zmap = zmap; k.ws = ws; k.ws_bytes = n;
    return std::string((const char*)&k, sizeof k);
}

static constexpr size_t kGraphCacheEntries = 256;

extern "C" {

int tsp_project_frame(tsp_handle* h, const tsp_frame_desc* desc, const uint16_t* d_stack, float* d_proj,
                      int32_t* d_zmap, void* d_workspace, size_t workspace_bytes, void* cuda_stream) {
    if (!h || !desc || !d_stack || !d_proj || !d_zmap || !d_workspace) {
        set_error("null argument");
        return TSP_ERR_INVALID;
    }
    TSP_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)cuda_stream;
    if (h->profiling || !h->dbg.graphs)
        return project_frame_eager(h, desc, d_stack, d_proj, d_zmap, d_workspace, workspace_bytes, s);
    const std::string key = graph_key(desc, d_stack, d_proj, d_zmap, d_workspace, workspace_bytes);
    std::unique_lock<std::mutex> lock(h->graph_mu);
    tsp_handle::GraphEntry* e = nullptr;
    for (auto& g : h->graphs)
        if (g.key == key) { e = &g; break; }
    if (!e) {
        if (h->graphs.size() >= kGraphCacheEntries) {              // evict the least recently used
            size_t victim = 0;
            for (size_t i = 1; i < h->graphs.size(); ++i)
                if (h->graphs[i].last_use < h->graphs[victim].last_use) victim = i;
            if (h->graphs[victim].exec) cudaGraphExecDestroy(h->graphs[victim].exec);
            h->graphs.erase(h->graphs.begin() + victim);
        }
        h->graphs.emplace_back();
        e = &h->graphs.back();
        e->key = key;
        e->last_use = ++h->graph_tick;
        lock.unlock();                                              // first sight: plain launches (tables get uploaded)
        return project_frame_eager(h, desc, d_stack, d_proj, d_zmap, d_workspace, workspace_bytes, s);
    }
    e->last_use = ++h->graph_tick;
    if (!e->exec && !e->failed) {
        if (!h->capture_stream) TSP_CUDA(cudaStreamCreateWithFlags(&h->capture_stream, cudaStreamNonBlocking));
        const int64_t before = h->launches;
        cudaGraph_t graph = nullptr;
        int rc = TSP_OK;
        if (cudaStreamBeginCapture(h->capture_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            rc = project_frame_eager(h, desc, d_stack, d_proj, d_zmap, d_workspace, workspace_bytes, h->capture_stream);
            const cudaError_t ce = cudaStreamEndCapture(h->capture_stream, &graph);
            if (rc == TSP_OK && ce == cudaSuccess && graph &&
                cudaGraphInstantiate(&e->exec, graph, 0) == cudaSuccess) {
                e->launches = h->launches - before;
            } else {
                e->exec = nullptr;
                e->failed = true;
            }
            if (graph) cudaGraphDestroy(graph);
        } else {
            e->failed = true;
        }
        h->launches = before;
        cudaGetLastError();                                         // a failed capture leaves a sticky-looking error behind
        if (rc != TSP_OK && rc != TSP_ERR_CUDA) return rc;          // argument errors are the caller's, not the capture's
    }
    if (e->exec) {
        cudaGraphExec_t exec = e->exec;
        const int64_t n = e->launches;
        const cudaError_t le = cudaGraphLaunch(exec, s);
        if (le != cudaSuccess) {
            set_error("cudaGraphLaunch failed: %s", cudaGetErrorString(le));
            return TSP_ERR_CUDA;
        }
        h->launches += n;
        h->graph_replays++;
        return TSP_OK;
    }
    lock.unlock();
    return project_frame_eager(h, desc, d_stack, d_proj, d_zmap, d_workspace, workspace_bytes, s);
}

int tsp_set_profiling(tsp_handle* h, int enable) {
    if (!h) return TSP_ERR_INVALID;
    h->profiling = enable != 0;
    return TSP_OK;
}

int tsp_stage_count(void) { return STG_COUNT; }

const char* tsp_stage_name(int stage) { return stage >= 0 && stage < STG_COUNT ? kStageNames[stage] : ""; }

// Folds every recorded mark into per-stage totals (synchronises the device), returns them and optionally
// resets.  ms_out / count_out hold tsp_stage_count() entries.
int tsp_get_stage_times(tsp_handle* h, double* ms_out, int64_t* count_out, int reset) {
    if (!h || !ms_out || !count_out) return TSP_ERR_INVALID;
    TSP_CUDA(cudaSetDevice(h->device));
    TSP_CUDA(cudaDeviceSynchronize());
    for (size_t i = 0; i < h->prof_events.size(); ++i) {
        const int st = h->prof_stage[i];
        if (st >= 0 && i > 0) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, h->prof_events[i - 1], h->prof_events[i]) == cudaSuccess) {
                h->prof_ms[st] += ms;
                h->prof_count[st] += 1;
            }
        }
    }
    for (cudaEvent_t ev : h->prof_events) h->prof_pool.push_back(ev);
    h->prof_events.clear();
    h->prof_stage.clear();
    for (int i = 0; i < STG_COUNT; ++i) {
        ms_out[i] = h->prof_ms[i];
        count_out[i] = h->prof_count[i];
        if (reset) {
            h->prof_ms[i] = 0;
            h->prof_count[i] = 0;
        }
    }
    return TSP_OK;
}

static void fill_status(const int32_t* st, tsp_frame_status* out) {
    memset(out, 0, sizeof *out);
    out->band_index_error = st[ST_BAND_ERR];
    out->has_nonzero = st[ST_HAS_NONZERO];
    memcpy(&out->percentile95, &st[ST_P95_BITS], sizeof(float));
    out->zmap_min = st[ST_ZMIN];
    out->zmap_max = st[ST_ZMAX];
    out->nonzero_count = (int64_t)(((uint64_t)(uint32_t)st[ST_NZ_HI] << 32) | (uint32_t)st[ST_NZ_LO]);
    out->near_tie_pixels = st[ST_NEAR_TIE];
}

int tsp_get_frame_status(tsp_handle* h, const void* d_workspace, void* cuda_stream, tsp_frame_status* out) {
    if (!h || !d_workspace || !out) return TSP_ERR_INVALID;
    TSP_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)cuda_stream;
    int32_t st[kStatusWords];
    TSP_CUDA(cudaMemcpyAsync(st, d_workspace, sizeof st, cudaMemcpyDeviceToHost, s));
    TSP_CUDA(cudaStreamSynchronize(s));
    fill_status(st, out);
    if (out->band_index_error) {
        set_error("height map indexes past the cropped stack (reference raises IndexError)");
        return TSP_ERR_BAND_INDEX;
    }
    return TSP_OK;
}

static int ensure_scratch(tsp_handle* h, size_t bytes) {
    if (h->d_scratch_bytes >= bytes) return TSP_OK;
    if (h->d_scratch) TSP_CUDA(cudaFree(h->d_scratch));
    h->d_scratch = nullptr;
    h->d_scratch_bytes = 0;
    TSP_CUDA(cudaMalloc(&h->d_scratch, bytes));
    h->d_scratch_bytes = bytes;
    return TSP_OK;
}

static int ensure_slot(tsp_handle* h, int slot, size_t bytes) {
    tsp_handle::Slot& sl = h->slots[slot];
    if (!sl.stream) TSP_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    if (!sl.h_status) TSP_CUDA(cudaMallocHost((void**)&sl.h_status, kStatusWords * sizeof(int32_t)));
    if (sl.bytes < bytes) {
        if (sl.d_mem) {
            TSP_CUDA(cudaStreamSynchronize(sl.stream));
            TSP_CUDA(cudaFree(sl.d_mem));
        }
        sl.d_mem = nullptr;
        sl.bytes = 0;
        TSP_CUDA(cudaMalloc(&sl.d_mem, bytes));
        sl.bytes = bytes;
    }
    return TSP_OK;
}

// copy-in, operator, conversion, copy-out of one frame on slot `slot` (< TSP_MAX_SLOTS: the caller's slots; the last
// one belongs to tsp_project_frame_host)
static int submit_on_slot(tsp_handle* h, int slot, const tsp_frame_desc* desc, const uint16_t* h_stack, void* h_proj,
                          void* h_zmap) {
    Crop c;
    int rc = resolve_crop(desc, &c);
    if (rc) return rc;
    TSP_CUDA(cudaSetDevice(h->device));
    std::lock_guard<std::mutex> lock(h->host_mu);
    tsp_handle::Slot& sl = h->slots[slot];
    if (sl.busy.load()) {
        set_error("slot %d still has a frame in flight: call tsp_frame_wait first", slot);
        return TSP_ERR_INVALID;
    }
    const bool out16 = (desc->flags & TSP_FRAME_OUT_U16) != 0;
    const size_t plane = (size_t)desc->rows * desc->cols;
    const size_t nstack = (size_t)desc->channels * desc->planes * plane;
    const size_t nproj = (size_t)desc->channels * plane;
    const size_t ws = tsp_project_workspace_bytes(desc);
    if (ws == 0) return TSP_ERR_INVALID;
    const size_t proj_out = nproj * (out16 ? sizeof(uint16_t) : sizeof(double));
    const size_t zmap_out = plane * (out16 ? sizeof(uint16_t) : sizeof(int64_t));
    size_t off = 0;
    const size_t o_stack = off;  off += align_up(nstack * sizeof(uint16_t), 256);
    const size_t o_proj = off;   off += align_up(nproj * sizeof(float), 256);
    const size_t o_zmap = off;   off += align_up(plane * sizeof(int32_t), 256);
    const size_t o_projx = off;  off += align_up(proj_out, 256);
    const size_t o_zmapx = off;  off += align_up(zmap_out, 256);
    const size_t o_ws = off;     off += ws;
    rc = ensure_slot(h, slot, off);
    if (rc) return rc;
    char* base = (char*)sl.d_mem;
    cudaStream_t s = sl.stream;
    tsp_frame_desc d2 = *desc;
    d2.flags &= ~TSP_FRAME_OUT_U16;
    for (int k = 0; k <= TSP_MAX_SLOTS; ++k)
        if (k != slot && h->slots[k].busy.load()) d2.flags |= TSP_FRAME_CONCURRENT;      // other frames are in flight
    // From the first enqueued copy on, a failure must not return while the stream may still read the caller's
    // buffers: drain the slot's stream first.
    auto fail = [&](int code) {
        cudaStreamSynchronize(s);
        return code;
    };
    cudaError_t e = cudaSuccess;
    if (nstack * sizeof(uint16_t) >= ((size_t)4 << 20) && !host_pointer_is_pinned(h_stack)) {
        rc = staged_copy_in(h, sl, base + o_stack, (const char*)h_stack, nstack * sizeof(uint16_t), s);
        if (rc) return fail(rc);
    } else {
        e = cudaMemcpyAsync(base + o_stack, h_stack, nstack * sizeof(uint16_t), cudaMemcpyHostToDevice, s);
    }
    if (e != cudaSuccess) {
        set_error("copy-in failed: %s", cudaGetErrorString(e));
        return fail(TSP_ERR_CUDA);
    }
    rc = tsp_project_frame(h, &d2, (const uint16_t*)(base + o_stack), (float*)(base + o_proj),
                           (int32_t*)(base + o_zmap), base + o_ws, ws, s);
    if (rc) return fail(rc);
    {
        NvtxRange r("tsp/convert_copy_out");
        if (out16)
            rc = launch_narrow_outputs(h, (const float*)(base + o_proj), (const int32_t*)(base + o_zmap),
                                       (uint16_t*)(base + o_projx), (uint16_t*)(base + o_zmapx), nproj, plane, s);
        else
            rc = launch_widen_outputs(h, (const float*)(base + o_proj), (const int32_t*)(base + o_zmap),
                                      (double*)(base + o_projx), (int64_t*)(base + o_zmapx), nproj, plane, s);
        if (rc) return fail(rc);
        e = cudaMemcpyAsync(sl.h_status, base + o_ws, kStatusWords * sizeof(int32_t), cudaMemcpyDeviceToHost, s);
        // results into pageable arrays (the ctypes stub of INTEGRATION.md allocates them with np.empty): through a
        // pinned block of the slot, handed over by the host threads in tsp_frame_wait - the driver's own staging of a
        // pageable destination runs at ~10 GB/s on one thread
        void* dst_proj = h_proj;
        void* dst_zmap = h_zmap;
        sl.user_proj = sl.user_zmap = nullptr;
        if (proj_out + zmap_out >= ((size_t)1 << 20) && !(host_pointer_is_pinned(h_proj) && host_pointer_is_pinned(h_zmap))) {
            const size_t need = align_up(proj_out, 256) + zmap_out;
            if (sl.h_out_bytes < need) {
                if (sl.h_out) cudaFreeHost(sl.h_out);
                sl.h_out = nullptr;
                sl.h_out_bytes = 0;
                if (cudaMallocHost(&sl.h_out, need) == cudaSuccess) sl.h_out_bytes = need;
                else cudaGetLastError();
            }
            if (sl.h_out) {
                if (!h->copy_pool) h->copy_pool = new CopyPool(copy_pool_threads());
                dst_proj = sl.h_out;
                dst_zmap = (char*)sl.h_out + align_up(proj_out, 256);
                sl.user_proj = h_proj;
                sl.user_zmap = h_zmap;
                sl.user_proj_bytes = proj_out;
                sl.user_zmap_bytes = zmap_out;
            }
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(dst_proj, base + o_projx, proj_out, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dst_zmap, base + o_zmapx, zmap_out, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) {
            set_error("copy-out failed: %s", cudaGetErrorString(e));
            return fail(TSP_ERR_CUDA);
        }
    }
    sl.busy.store(true);
    return TSP_OK;
}

static int wait_on_slot(tsp_handle* h, int slot, tsp_frame_status* status) {
    tsp_handle::Slot& sl = h->slots[slot];
    if (!sl.busy.load()) {
        set_error("slot %d has no frame in flight", slot);
        return TSP_ERR_INVALID;
    }
    TSP_CUDA(cudaSetDevice(h->device));
    cudaError_t e = cudaStreamSynchronize(sl.stream);
    tsp_frame_status st;
    if (e == cudaSuccess) fill_status(sl.h_status, &st);
    if (sl.user_proj) {                       // results staged in the slot's pinned block: hand them over
        if (e == cudaSuccess && h->copy_pool) {
            h->copy_pool->copy(sl.user_proj, sl.h_out, sl.user_proj_bytes);
            h->copy_pool->copy(sl.user_zmap, (char*)sl.h_out + align_up(sl.user_proj_bytes, 256), sl.user_zmap_bytes);
        }
        sl.user_proj = sl.user_zmap = nullptr;
    }
    sl.busy.store(false);
    if (e != cudaSuccess) {
        set_error("cudaStreamSynchronize failed: %s", cudaGetErrorString(e));
        return TSP_ERR_CUDA;
    }
    if (status) *status = st;
    if (st.band_index_error) {
        set_error("height map indexes past the cropped stack (reference raises IndexError)");
        return TSP_ERR_BAND_INDEX;
    }
    return TSP_OK;
}

int tsp_frame_submit(tsp_handle* h, int slot, const tsp_frame_desc* desc, const uint16_t* h_stack, void* h_proj,
                     void* h_zmap) {
    if (!h || !desc || !h_stack || !h_proj || !h_zmap || slot < 0 || slot >= TSP_MAX_SLOTS) {
        set_error("bad argument to tsp_frame_submit");
        return TSP_ERR_INVALID;
    }
    return submit_on_slot(h, slot, desc, h_stack, h_proj, h_zmap);
}

int tsp_frame_wait(tsp_handle* h, int slot, tsp_frame_status* status) {
    if (!h || slot < 0 || slot >= TSP_MAX_SLOTS) return TSP_ERR_INVALID;
    return wait_on_slot(h, slot, status);
}

// One blocking call = submit + wait on the handle's own slot; single_mu makes calls from several host threads take
// turns (a slot pipeline running on the same handle is not disturbed: its slots are different ones).
int tsp_project_frame_host(tsp_handle* h, const tsp_frame_desc* desc, const uint16_t* h_stack, void* h_proj,
                           void* h_zmap, tsp_frame_status* status) {
    if (!h || !desc || !h_stack || !h_proj || !h_zmap) {
        set_error("bad argument to tsp_project_frame_host");
        return TSP_ERR_INVALID;
    }
    std::lock_guard<std::mutex> turn(h->single_mu);
    int rc = submit_on_slot(h, TSP_MAX_SLOTS, desc, h_stack, h_proj, h_zmap);
    if (rc) return rc;
    return wait_on_slot(h, TSP_MAX_SLOTS, status);
}

int tsp_gaussian_blur_f32(tsp_handle* h, const float* d_in, float* d_out, float* d_tmp, int planes, int rows,
                          int cols, const double sigma[3], int fp64_accumulate, void* cuda_stream) {
    if (!h || !d_in || !d_out || !d_tmp || !sigma || planes < 1 || rows < 1 || cols < 1) return TSP_ERR_INVALID;
    TSP_CUDA(cudaSetDevice(h->device));
    return gaussian_blur<float>(h, d_in, d_out, d_tmp, planes, rows, cols, sigma, fp64_accumulate != 0,
                                (cudaStream_t)cuda_stream);
}

int tsp_gaussian_blur_u16(tsp_handle* h, const uint16_t* d_in, uint16_t* d_out, uint16_t* d_tmp, int planes,
                          int rows, int cols, const double sigma[3], void* cuda_stream) {
    if (!h || !d_in || !d_out || !d_tmp || !sigma || planes < 1 || rows < 1 || cols < 1) return TSP_ERR_INVALID;
    TSP_CUDA(cudaSetDevice(h->device));
    return gaussian_blur<uint16_t>(h, d_in, d_out, d_tmp, planes, rows, cols, sigma, true,
                                   (cudaStream_t)cuda_stream);
}

int tsp_percentile95_nonzero_u16(tsp_handle* h, const uint16_t* d_volume, size_t count, int airyscan,
                                 void* cuda_stream, tsp_frame_status* out) {
    return tsp_percentile_nonzero_u16(h, d_volume, count, 95.0f, airyscan ? kAiryscanPedestal : 0, cuda_stream, out);
}

int tsp_percentile_nonzero_u16(tsp_handle* h, const uint16_t* d_volume, size_t count, float percentile, int pedestal,
                               void* cuda_stream, tsp_frame_status* out) {
    if (!h || !d_volume || !out) return TSP_ERR_INVALID;
    if (!(percentile >= 0.0f && percentile <= 100.0f) || pedestal < 0 || pedestal > 65535) {
        set_error("percentile %g / pedestal %d out of range", (double)percentile, pedestal);
        return TSP_ERR_INVALID;
    }
    TSP_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)cuda_stream;
    const size_t bytes = align_up(kStatusWords * sizeof(int32_t), 256) + percentile_scratch_bytes();
    std::lock_guard<std::mutex> lock(h->host_mu);
    int rc = ensure_scratch(h, bytes);
    if (rc) return rc;
    int32_t* st = (int32_t*)h->d_scratch;
    uint32_t* hist = (uint32_t*)((char*)h->d_scratch + align_up(kStatusWords * sizeof(int32_t), 256));
    rc = launch_percentile(h, d_volume, count, pedestal, st, hist, s, nullptr, 0, percentile / 100.0f,
                           (double)percentile / 100.0);
    if (rc) return rc;
    TSP_CUDA(cudaMemcpyAsync(h->h_status, st, kStatusWords * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    TSP_CUDA(cudaStreamSynchronize(s));
    fill_status(h->h_status, out);
    return TSP_OK;
}

int tsp_focus_score_f32(tsp_handle* h, const uint16_t* d_channel, float* d_score, float* d_tmp, int planes,
                        int rows, int cols, int airyscan, int fp64_accumulate, void* cuda_stream) {
    if (!h || !d_channel || !d_score || !d_tmp || planes < 1 || rows < 1 || cols < 1) return TSP_ERR_INVALID;
    TSP_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)cuda_stream;
    const size_t bytes = align_up(kStatusWords * sizeof(int32_t), 256) + percentile_scratch_bytes();
    std::lock_guard<std::mutex> lock(h->host_mu);
    int rc = ensure_scratch(h, bytes);
    if (rc) return rc;
    int32_t* st = (int32_t*)h->d_scratch;
    uint32_t* hist = (uint32_t*)((char*)h->d_scratch + align_up(kStatusWords * sizeof(int32_t), 256));
    const size_t nvox = (size_t)planes * rows * cols;
    const int ped = airyscan ? kAiryscanPedestal : 0;
    rc = launch_percentile(h, d_channel, nvox, ped, st, hist, s);
    if (rc) return rc;
    rc = launch_prepare(h, d_channel, d_score, nvox, ped, st, s);
    if (rc) return rc;
    const double sig_pre[3] = {0.5, 1.0, 1.0}, sig_score[3] = {0.5, 30.0, 30.0};
    const bool fp64 = fp64_accumulate != 0;
    rc = gaussian_blur<float>(h, d_score, d_tmp, d_score, planes, rows, cols, sig_pre, fp64, s);
    if (rc) return rc;
    return gaussian_blur<float>(h, d_tmp, d_score, d_tmp, planes, rows, cols, sig_score, fp64, s);
}

int tsp_argmax_z_f32(tsp_handle* h, const float* d_score, int32_t* d_zmap, int planes, int rows, int cols,
                     int z_offset, void* cuda_stream) {
    if (!h || !d_score || !d_zmap || planes < 1 || rows < 1 || cols < 1) return TSP_ERR_INVALID;
    TSP_CUDA(cudaSetDevice(h->device));
    return launch_argmax(h, d_score, d_zmap, planes, rows, cols, z_offset, nullptr, (cudaStream_t)cuda_stream);
}

size_t tsp_manifold_workspace_bytes(void) {
    return align_up(kStatusWords * sizeof(int32_t), 256) + align_up(manifold_scratch_bytes(), 256);
}

int tsp_build_manifold(tsp_handle* h, const float* d_score, int32_t* d_chosen, int planes, int rows, int cols,
                       void* d_workspace, size_t workspace_bytes, void* cuda_stream) {
    if (!h || !d_score || !d_chosen || !d_workspace || planes < 1 || rows < 1 || cols < 1) return TSP_ERR_INVALID;
    if (workspace_bytes < tsp_manifold_workspace_bytes()) return TSP_ERR_WORKSPACE;
    TSP_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)cuda_stream;
    TSP_CUDA(cudaMemsetAsync(d_workspace, 0, kStatusWords * sizeof(int32_t), s));
    return launch_manifold(h, d_score, d_chosen, planes, rows, cols, (int32_t*)d_workspace,
                           (char*)d_workspace + align_up(kStatusWords * sizeof(int32_t), 256), s);
}

int tsp_block_reduce_f32(tsp_handle* h, const float* d_volume, float* d_out, int planes, int rows, int cols,
                         int bin_size, int variance, void* cuda_stream) {
    if (!h || !d_volume || !d_out || planes < 1 || rows < 1 || cols < 1 || bin_size < 1) return TSP_ERR_INVALID;
    TSP_CUDA(cudaSetDevice(h->device));
    return launch_block_reduce(h, d_volume, d_out, planes, rows, cols, bin_size, variance != 0, false,
                               (cudaStream_t)cuda_stream);
}

int tsp_resize_argmax_f32(tsp_handle* h, const float* d_score, int32_t* d_zmap, int planes, int rows, int cols,
                          int coarse_rows, int coarse_cols, int z_offset, void* d_workspace, size_t workspace_bytes,
                          void* cuda_stream) {
    if (!h || !d_score || !d_zmap || !d_workspace || planes < 1 || rows < 1 || cols < 1 || coarse_rows < 1 ||
        coarse_cols < 1)
        return TSP_ERR_INVALID;
    if (workspace_bytes < kStatusWords * sizeof(int32_t)) return TSP_ERR_WORKSPACE;
    TSP_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)cuda_stream;
    TSP_CUDA(cudaMemsetAsync(d_workspace, 0, kStatusWords * sizeof(int32_t), s));
    return launch_resize_argmax(h, d_score, d_zmap, planes, rows, cols, coarse_rows, coarse_cols, z_offset,
                                (int32_t*)d_workspace, s);
}

int tsp_resize_round_i32(tsp_handle* h, const int32_t* d_coarse, int32_t* d_zmap, int rows, int cols,
                         int coarse_rows, int coarse_cols, void* cuda_stream) {
    if (!h || !d_coarse || !d_zmap || rows < 1 || cols < 1 || coarse_rows < 1 || coarse_cols < 1) return TSP_ERR_INVALID;
    TSP_CUDA(cudaSetDevice(h->device));
    return launch_resize_round(h, d_coarse, d_zmap, rows, cols, coarse_rows, coarse_cols, 0, 0,
                               (cudaStream_t)cuda_stream);
}

// [status block][worklist of the tiles the shallow-range band kernel leaves to the generic one]
size_t tsp_band_workspace_bytes(int, int, int rows, int cols) {
    return align_up(kStatusWords * sizeof(int32_t), 256) + align_up(band_worklist_bytes(rows, cols), 256);
}

int tsp_band_project(tsp_handle* h, const uint16_t* d_stack, const int32_t* d_zmap, float* d_proj, int channels,
                     int planes, int rows, int cols, int reference_channel, int atoh_shift, int airyscan,
                     void* d_workspace, size_t workspace_bytes, void* cuda_stream) {
    if (!h || !d_stack || !d_zmap || !d_proj || !d_workspace || channels < 1 || planes < 1 || rows < 1 || cols < 1)
        return TSP_ERR_INVALID;
    if (workspace_bytes < tsp_band_workspace_bytes(channels, planes, rows, cols)) return TSP_ERR_WORKSPACE;
    TSP_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)cuda_stream;
    TSP_CUDA(cudaMemsetAsync(d_workspace, 0, kStatusWords * sizeof(int32_t), s));
    const size_t plane = (size_t)rows * cols;
    struct PlainLaunches {               // the first kernel follows a memset: no programmatic launches for this call
        PlainLaunches() { tl_chain_launches = false; }
        ~PlainLaunches() { tl_chain_launches = true; }
    } plain;
    return launch_band_project_ex(h, d_stack, (size_t)planes * plane, 0, d_zmap, d_proj, channels, planes, rows,
                                  cols, reference_channel, atoh_shift, airyscan ? kAiryscanPedestal : 0,
                                  (int32_t*)d_workspace, false, s,
                                  (int*)((char*)d_workspace + align_up(kStatusWords * sizeof(int32_t), 256)));
}

size_t tsp_project_m_workspace_bytes(int planes, int rows, int cols, int bin_size) {
    if (planes < 1 || rows < 1 || cols < 1 || bin_size < 1) return 0;
    const size_t vox = align_up((size_t)planes * rows * cols, 128);
    const size_t by = (rows + bin_size - 1) / bin_size, bx = (cols + bin_size - 1) / bin_size;
    return 2 * vox * sizeof(uint16_t) + (size_t)planes * by * bx * sizeof(double) + 256;
}

int tsp_project_m(tsp_handle* h, const uint16_t* d_channel, uint16_t* d_out, int planes, int rows, int cols,
                  int method, int bin_size, void* d_workspace, size_t workspace_bytes, void* cuda_stream) {
    if (!h || !d_channel || !d_out || !d_workspace || planes < 1 || rows < 1 || cols < 1 || bin_size < 1 ||
        method < 0 || method > 1)
        return TSP_ERR_INVALID;
    if (planes >= 64) {
        set_error("np.choose accepts fewer than 64 planes (reference raises ValueError)");
        return TSP_ERR_CHOOSE_LIMIT;
    }
    if (workspace_bytes < tsp_project_m_workspace_bytes(planes, rows, cols, bin_size)) return TSP_ERR_WORKSPACE;
    TSP_CUDA(cudaSetDevice(h->device));
    return launch_project_m(h, d_channel, d_out, planes, rows, cols, method, bin_size, d_workspace,
                            (cudaStream_t)cuda_stream);
}

}  // extern "C"
