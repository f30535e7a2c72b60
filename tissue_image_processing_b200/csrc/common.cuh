// Shared declarations of libtsp_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <cstdlib>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/tsp_b200.h"

namespace tsp {

constexpr int kAiryscanPedestal = 10000;   // SP:28
constexpr int kHistBins = 65536;
constexpr int kStatusWords = 64;           // device status block (int32 words) at workspace start

// indices into the device status block
enum StatusWord {
    ST_BAND_ERR = 0,
    ST_HAS_NONZERO = 1,
    ST_P95_BITS = 2,
    ST_ZMIN = 3,
    ST_ZMAX = 4,
    ST_NZ_LO = 5,
    ST_NZ_HI = 6,
    ST_NEAR_TIE = 7,
    ST_WIN_LO = 8,       // percentile: first raw value of the counted window
    ST_WIN_N = 9,        //             window width in values
    ST_WIN_OK = 10,      //             1 when the sample produced a usable window
    ST_NEED_FULL = 11,   //             1 arms the full-histogram fallback
    ST_ZMIN_INV = 12,    // max over pixels of INT_MAX - z (so that a zeroed block is the identity)
    ST_WORK_COUNT = 13,  // band stage: tiles on the deep-range worklist
};

void set_error(const char* fmt, ...);

// tsp_params with the defaults applied and the routing decisions taken (api.cu: resolve_params)
struct Params {
    float q;                  // percentile / 100 in float32, the way numpy forms it for float32 data (SP:35)
    double q64;               // the same fraction for the sample-window statistics
    int pedestal;             // 0 when the frame is not airyscan
    double sigma_pre[3], sigma_score[3], sigma_mask[3];
    bool default_score;       // sigma_pre / sigma_score are the reference's: the fast-mode tables apply
    bool default_mask;        // sigma_mask is the reference's: the fused band kernels apply
};

enum Stage {
    STG_PERCENTILE = 0, STG_DECIMATE, STG_COARSE, STG_INTERP_ARGMAX, STG_PREPARE, STG_BLUR_PRE, STG_BLUR_SCORE,
    STG_ARGMAX, STG_BAND, STG_WIDEN, STG_PCT_SAMPLE, STG_PCT_COUNT, STG_COUNT
};

#define TSP_CUDA(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            tsp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                           __LINE__);                                                        \
            return TSP_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)

#define TSP_LAUNCH_CHECK(h)                                                                    \
    do {                                                                                       \
        (h)->launches++;                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e == cudaSuccess && tsp::g_sync_launches) _e = cudaDeviceSynchronize();           \
        if (_e != cudaSuccess) {                                                               \
            tsp::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, \
                           __LINE__);                                                          \
            return TSP_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// NVTX range on the launching thread around the launches of one stage ("tsp/decimate", ...): a timeline shows which
// kernels belong to which stage of which frame.  Costs nothing without a profiler attached.
struct NvtxRange {
    bool open = true;
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    void end() { if (open) { nvtxRangePop(); open = false; } }
    ~NvtxRange() { end(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// scipy.ndimage Gaussian weights: radius int(4*sigma+0.5), exp(-k^2/(2 sigma^2)) / sum, float64
std::vector<double> gaussian_taps(double sigma);

struct DeviceTaps {          // one uploaded FIR: w64[0..2r], w32 padded with `pad` zeros each side
    int radius = 0;
    const double* w64 = nullptr;
    const float* w32 = nullptr;    // points at the first real tap; w32[-pad..-1] and [2r+1..2r+pad] are 0
};
constexpr int kTapPad = 8;

// ---- programmatic dependent launch (the kernels of one frame form a chain on one stream) ----------------
// A kernel launched with launch_chained() may be scheduled while its predecessor on the stream is still draining:
// its CTAs take the SM slots the predecessor's last CTAs free, run whatever does not touch global memory (shared
// memory clears, mbarrier setup, index arithmetic) and then block in chain_wait() until the predecessor grid has
// completed and its writes are visible.  Every kernel calls chain_release() first thing so that its own successor
// may start queueing.  Launched the ordinary way (<<<>>>) both calls do nothing.  What this buys is the ~2 us of
// launch latency and ramp at each of the frame's ten kernel boundaries.
extern thread_local bool tl_chain_launches;      // false while a TSP_FRAME_CONCURRENT frame is being enqueued
extern std::atomic<bool> g_no_chain;
extern std::atomic<int> g_chain_plain_mask;      // debugging aid (tsp_debug_set "chain_plain_mask"): bit i = the i-th chained launch of a frame is a plain one
extern thread_local int tl_chain_site;           // chained launches seen since the frame started
extern bool g_sync_launches;                     // debugging aid (TSP_SYNC_LAUNCHES=1): device sync + error check per launch             // A/B switch (TSP_NO_CHAIN at tsp_create, tsp_debug_set "no_chain")
#ifdef __CUDACC__
__device__ __forceinline__ void chain_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// A chained kernel is resident - and the L1 flush of its launch long past - while its predecessors are still running,
// and CTAs of those predecessors share the SM.  griddepcontrol.wait makes the predecessor's writes visible at L2, it
// does NOT drop L1 lines this SM loaded earlier: a line read (through L1) before the producer's last write is served
// stale afterwards.  Found with a 30 x 1024 x 1024 stack: from the second frame on coarse_xy_kernel took the
// percentile from a status line cached before window_count_kernel wrote it (fixed-point scale of an un-clipped frame):
// every height map wrong, deterministically.  Rules for chained kernels:
//   * words of the status block are read with st_load() (ld.global.cg: L2, never L1);
//   * memory that is re-written within the frame after an earlier stage read it (the accumulation volume, whose
//     memory the z-mixed volume re-uses) is read with __ldcg by the earlier stage, so no L1 line of it exists;
//   * everything else a stage reads was last read before the frame's first kernel, which is launched the ordinary way
//     (a full boundary: L1 dropped).
// (A gpu-scope fence behind the wait - CCTL.IVALL - also cures it, at +19 us in the decimation kernel, whose 9 728
// warps each pass the wait.)
__device__ __forceinline__ void chain_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ int32_t st_load(const int32_t* status, int word) { return __ldcg(status + word); }

// launch_after_copy(): the same launch WITHOUT the attribute, for a kernel whose predecessor on the stream is a memset
// or copy rather than a kernel: there is no prologue to overlap with, and inside a captured graph an early start
// against a non-kernel node is not something to rely on.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_after_copy(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                     Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cfg.attrs = nullptr;
    cfg.numAttrs = 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                  Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    const int site = tl_chain_site++;
    const bool plain = site < 31 && ((g_chain_plain_mask.load(std::memory_order_relaxed) >> site) & 1);
    cfg.numAttrs = !g_no_chain.load(std::memory_order_relaxed) && tl_chain_launches && !plain ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---- TMA / mbarrier primitives (sm_100a) ----------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
// 1-D bulk copy global -> shared (bytes a multiple of 16, both sides 16-byte aligned), completing on an mbarrier
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int z, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar)
        : "memory");
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int x, int y, int z, int c,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
        "l"(map), "r"(x), "r"(y), "r"(z), "r"(c), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(x), "r"(y), "r"(bar)
        : "memory");
}
#endif

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int get_tensor_map_encoder(EncodeTiledFn* out);

struct CopyPool;           // host threads of the staged copy-in (api.cu)
}  // namespace tsp

struct tsp_handle {
    int device = 0;
    int sm_count = 148;
    int64_t launches = 0;
    // kernel-variant switches for tests / A-B measurements: environment read once at tsp_create, tsp_debug_set later
    struct Debug {
        int no_ring = 0;         // strip decimation kernel instead of the TMA ring
        int band_variant = 0;    // 0 auto, 2 register-prefetch kernel, 3 TMA ring kernel for every tile
        int interp_rows = 4;     // image rows per thread of the interpolation + argmax stage
        int graphs = 1;          // replay a frame's launch sequence as a CUDA graph (api.cu)
        int interp_global = 0;   // interpolation stage reads its control points from L2 (round-1 kernel) instead of shared memory
    } dbg;
    // CUDA graphs of frames seen before, keyed by descriptor + buffer pointers (api.cu: tsp_project_frame)
    struct GraphEntry {
        std::string key;
        cudaGraphExec_t exec = nullptr;
        int64_t launches = 0;
        uint64_t last_use = 0;
        bool failed = false;
    };
    std::vector<GraphEntry> graphs;
    std::mutex graph_mu;
    cudaStream_t capture_stream = nullptr;
    uint64_t graph_tick = 0;
    int64_t graph_replays = 0;
    std::mutex mu;        // guards taps / tables
    std::mutex host_mu;   // guards d_scratch and the slot table (held only while a call is being enqueued)
    std::mutex single_mu; // tsp_project_frame_host: one blocking call at a time on the handle's own slot
    // small device arena for FIR taps and fast-mode tables, keyed by a text key
    std::map<std::string, tsp::DeviceTaps> taps;
    std::map<std::string, void*> tables;
    // host-buffer path state (owned scratch, grows on demand)
    cudaStream_t stream = nullptr;
    void* d_scratch = nullptr;
    size_t d_scratch_bytes = 0;
    int32_t* h_status = nullptr;   // pinned copy of the status block
    // host-buffer frame slots (tsp_frame_submit / tsp_frame_wait): own stream + device memory each, so
    // the H2D copy of one frame overlaps the kernels of another and the D2H copy of a third
    struct Slot {
        cudaStream_t stream = nullptr;
        void* d_mem = nullptr;
        size_t bytes = 0;
        int32_t* h_status = nullptr;
        std::atomic<bool> busy{false};
        // pageable host buffers: ring of pinned chunks the stack is staged through (api.cu: staged_copy_in),
        // allocated on first use
        static constexpr int kStageRing = 3;
        void* stage[kStageRing] = {nullptr, nullptr, nullptr};
        cudaEvent_t stage_ev[kStageRing] = {nullptr, nullptr, nullptr};
        // pageable result buffers: the frame's outputs land in this pinned block and are copied to the caller's
        // arrays by the host threads when the frame is waited for
        void* h_out = nullptr;
        size_t h_out_bytes = 0;
        void* user_proj = nullptr;
        void* user_zmap = nullptr;
        size_t user_proj_bytes = 0, user_zmap_bytes = 0;
    };
    tsp::CopyPool* copy_pool = nullptr;        // host threads of the staged copies (api.cu), created on first use
    Slot slots[TSP_MAX_SLOTS + 1];       // the last one belongs to tsp_project_frame_host
    // per-device one-time setup done (constant memory, function attributes)
    size_t interp_smem_set = 0;
    bool fast_consts = false, band_consts = false, hist_attr = false, ring_attr = false, band3_attr = false, band4_attr = false, xy_attr = false, manifold_attr = false, count_attr = false;
    // optional per-stage timing (tsp_set_profiling): CUDA events recorded on the launching stream
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;       // marks of calls not yet folded into the totals
    std::vector<int> prof_stage;                // stage id attributed to the interval ending at the mark (-1 = start)
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[16] = {0};
    int64_t prof_count[16] = {0};
};

namespace tsp {

int get_taps(tsp_handle* h, double sigma, DeviceTaps* out);
void prof_mark(tsp_handle* h, cudaStream_t s, int stage);   // stage -1 opens a timed sequence

// stage launchers (each returns TSP_OK or an error code)
size_t percentile_scratch_bytes();
int launch_percentile(tsp_handle* h, const uint16_t* d_vol, size_t count, int pedestal, int32_t* d_status,
                      void* d_scratch, cudaStream_t s, void* zero_ptr = nullptr, size_t zero_bytes = 0,
                      float q = 0.95f, double q64 = 0.95);
int launch_prepare(tsp_handle* h, const uint16_t* d_in, float* d_out, size_t count, int pedestal,
                   const int32_t* d_status, cudaStream_t s);
int launch_convert_u16_f32(tsp_handle* h, const uint16_t* d_in, float* d_out, size_t count,
                           cudaStream_t s);
template <typename T>
int launch_fir_axis(tsp_handle* h, const T* d_in, T* d_out, int Z, int Y, int X, int axis,
                    const DeviceTaps& taps, bool fp64, cudaStream_t s);
template <typename T>
int gaussian_blur(tsp_handle* h, const T* d_in, T* d_out, T* d_tmp, int Z, int Y, int X,
                  const double sigma[3], bool fp64, cudaStream_t s);
int prepare_and_blur_f32(tsp_handle* h, const uint16_t* d_in, float* d_out, float* d_tmp, int Z, int Y, int X,
                         const double sigma[3], int pedestal, const int32_t* d_status, cudaStream_t s);
int launch_argmax(tsp_handle* h, const float* d_score, int32_t* d_zmap, int Z, int Y, int X,
                  int z_offset, int32_t* d_status, cudaStream_t s);
int launch_band_project_ex(tsp_handle* h, const uint16_t* d_stack, size_t channel_stride, size_t z0_offset,
                           const int32_t* d_zmap, float* d_proj, int C, int Z, int Y, int X, int ref_c,
                           int shift, int pedestal, int32_t* d_status, bool range_known, cudaStream_t s,
                           int* d_worklist = nullptr, const int32_t* d_zmap_other = nullptr);
size_t band_worklist_bytes(int Y, int X);
// binned scores, their resampling and the continuous manifold (binned.cu)
int launch_percentile_all(tsp_handle* h, const uint16_t* d_vol, size_t count, int pedestal, int32_t* d_status,
                          void* d_scratch, cudaStream_t s, float q = 0.95f);
int launch_block_reduce(tsp_handle* h, const float* d_vol, float* d_out, int Z, int Y, int X, int bin,
                        bool variance, bool multiply, cudaStream_t s);
int launch_resize_argmax(tsp_handle* h, const float* d_score, int32_t* d_zmap, int Z, int Y, int X, int cy, int cx,
                         int z_offset, int32_t* d_status, cudaStream_t s);
int launch_resize_round(tsp_handle* h, const int32_t* d_coarse, int32_t* d_zmap, int Y, int X, int cy, int cx,
                        int shift, int clip_hi, cudaStream_t s);
size_t manifold_scratch_bytes();
int launch_manifold(tsp_handle* h, const float* d_score, int32_t* d_chosen, int P, int R, int C, int32_t* d_status,
                    void* d_scratch, cudaStream_t s);
int launch_band_project_bitexact_ex(tsp_handle* h, const uint16_t* d_stack, size_t channel_stride,
                                    size_t z0_offset, const int32_t* d_zmap, float* d_proj, int C, int Z,
                                    int Y, int X, int ref_c, int shift, int pedestal, float* d_volA,
                                    float* d_volB, int32_t* d_status, bool range_known, cudaStream_t s,
                                    const int32_t* d_zmap_other = nullptr, const double* sigma_mask = nullptr,
                                    bool fp64 = true);
int launch_project_m(tsp_handle* h, const uint16_t* d_channel, uint16_t* d_out, int Z, int Y, int X,
                     int method, int bin, void* d_ws, cudaStream_t s);
int launch_widen_outputs(tsp_handle* h, const float* d_proj, const int32_t* d_zmap, double* d_proj64,
                         int64_t* d_zmap64, size_t nproj, size_t nz, cudaStream_t s);
int launch_narrow_outputs(tsp_handle* h, const float* d_proj, const int32_t* d_zmap, uint16_t* d_proj16,
                          uint16_t* d_zmap16, size_t nproj, size_t nz, cudaStream_t s);

// fast (multirate) score stage
size_t fast_workspace_bytes(int Z, int Y, int X);
size_t fast_accum_bytes(int Z, int Y, int X);      // leading bytes of the fast workspace that must be zero on entry
int launch_fast_score_argmax(tsp_handle* h, const uint16_t* d_channel, int32_t* d_zmap, int Z, int Y,
                             int X, int pedestal, int z_offset, int32_t* d_status, void* d_ws,
                             cudaStream_t s, bool accum_zeroed = false);

}  // namespace tsp
