// Stage K1 - 95th percentile of the non-zero voxels (reference surface_projection.py:32-36),
// exact: a full 65536-bin histogram of the raw uint16 values, then numpy's 'linear' percentile
// index arithmetic replayed in float32 (numpy/lib/_function_base_impl.py: virtual index
// (n-1)*q, floor, +1, gamma, _lerp - all float32 for float32 input; SURVEY trap T1).
#include "common.cuh"

namespace tsp {

// ---- histogram --------------------------------------------------------------------------------
// Per-CTA privatised histogram in shared memory with two 16-bit counters per 32-bit word
// (65536 bins = 128 KB).  A CTA adds at most kChunk < 65536 voxels between flushes, so no
// counter can overflow; a flush adds the non-zero counters to the global histogram.
constexpr int kHistThreads = 1024;
constexpr int kVecPerThread = 7;                                      // uint4 loads per thread per chunk
static_assert(kHistThreads * kVecPerThread * 8 + 16 < 65536, "16-bit counters must not overflow between flushes");
constexpr int kHistSmemBytes = kHistBins / 2 * 4;                     // 131072

__device__ __forceinline__ void hist_add(uint32_t* sh, uint32_t v) {
    atomicAdd(&sh[v >> 1], (v & 1u) ? 0x10000u : 1u);
}

__device__ __forceinline__ void hist_add_word(uint32_t* sh, uint32_t w) {
    hist_add(sh, w & 0xffffu);
    hist_add(sh, w >> 16);
}

__global__ void __launch_bounds__(kHistThreads, 1)
hist16_kernel(const uint16_t* __restrict__ vol, size_t count, uint32_t* __restrict__ ghist) {
    extern __shared__ uint32_t sh[];
    for (int i = threadIdx.x; i < kHistBins / 2; i += kHistThreads) sh[i] = 0;
    __syncthreads();

    // split [0,count) into a scalar head up to 16-byte alignment, a vector body, a scalar tail
    const uintptr_t addr = reinterpret_cast<uintptr_t>(vol);
    size_t head = ((16 - (addr & 15)) & 15) / 2;
    if (head > count) head = count;
    const size_t nvec = (count - head) / 8;
    const size_t tail0 = head + nvec * 8;
    const uint4* body = reinterpret_cast<const uint4*>(vol + head);

    if (blockIdx.x == 0) {
        for (size_t i = threadIdx.x; i < head; i += kHistThreads) hist_add(sh, vol[i]);
        for (size_t i = tail0 + threadIdx.x; i < count; i += kHistThreads) hist_add(sh, vol[i]);
    }

    const size_t vec_per_chunk = (size_t)kHistThreads * kVecPerThread;
    const size_t nchunks = (nvec + vec_per_chunk - 1) / vec_per_chunk;
    for (size_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const size_t base = chunk * vec_per_chunk;
        uint4 v[kVecPerThread];
#pragma unroll
        for (int j = 0; j < kVecPerThread; ++j) {
            const size_t idx = base + (size_t)j * kHistThreads + threadIdx.x;
            v[j] = idx < nvec ? __ldg(body + idx) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int j = 0; j < kVecPerThread; ++j) {
            const size_t idx = base + (size_t)j * kHistThreads + threadIdx.x;
            if (idx < nvec) {
                hist_add_word(sh, v[j].x);
                hist_add_word(sh, v[j].y);
                hist_add_word(sh, v[j].z);
                hist_add_word(sh, v[j].w);
            }
        }
        __syncthreads();
        // flush: 32 words per thread, vectorised
        uint4* sh4 = reinterpret_cast<uint4*>(sh);
        for (int i = threadIdx.x; i < kHistBins / 8; i += kHistThreads) {
            uint4 w = sh4[i];
            if (w.x | w.y | w.z | w.w) {
                const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t lo = ws[q] & 0xffffu, hi = ws[q] >> 16;
                    if (lo) atomicAdd(&ghist[(i * 4 + q) * 2], lo);
                    if (hi) atomicAdd(&ghist[(i * 4 + q) * 2 + 1], hi);
                }
                sh4[i] = make_uint4(0, 0, 0, 0);
            }
        }
        __syncthreads();
    }
    // head/tail voxels of block 0 when there was no chunk to flush them with
    if (blockIdx.x == 0) {
        __syncthreads();
        for (int i = threadIdx.x; i < kHistBins / 2; i += kHistThreads) {
            const uint32_t w = sh[i];
            if (w) {
                if (w & 0xffffu) atomicAdd(&ghist[i * 2], w & 0xffffu);
                if (w >> 16) atomicAdd(&ghist[i * 2 + 1], w >> 16);
            }
        }
    }
}

int launch_histogram(tsp_handle* h, const uint16_t* d_vol, size_t count, uint32_t* d_hist,
                     cudaStream_t s) {
    if (!h->hist_attr) {
        TSP_CUDA(cudaFuncSetAttribute(hist16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kHistSmemBytes));
        h->hist_attr = true;
    }
    TSP_CUDA(cudaMemsetAsync(d_hist, 0, kHistBins * sizeof(uint32_t), s));
    const size_t nchunks = (count / 8 + (size_t)kHistThreads * kVecPerThread - 1) /
                           ((size_t)kHistThreads * kVecPerThread);
    int grid = (int)(nchunks < (size_t)h->sm_count ? (nchunks ? nchunks : 1) : (size_t)h->sm_count);
    hist16_kernel<<<grid, kHistThreads, kHistSmemBytes, s>>>(d_vol, count, d_hist);
    TSP_LAUNCH_CHECK(h);
    return TSP_OK;
}

// ---- percentile from the histogram ----------------------------------------------------------------
// One CTA.  Bin b holds raw value b; after the optional pedestal the voxel value is b - pedestal,
// non-zero when b > pedestal.
__global__ void __launch_bounds__(1024, 1)
percentile_finalize_kernel(const uint32_t* __restrict__ ghist, int pedestal, int32_t* __restrict__ status) {
    __shared__ unsigned long long part[1024];
    __shared__ unsigned long long total_s;
    __shared__ int val_s[2];
    const int t = threadIdx.x;
    const int b0 = t * 64;
    unsigned long long mine = 0;
    for (int i = 0; i < 64; ++i) {
        const int b = b0 + i;
        if (b > pedestal) mine += ghist[b];
    }
    part[t] = mine;
    __syncthreads();
    // inclusive scan (Hillis-Steele on 1024 entries)
    for (int off = 1; off < 1024; off <<= 1) {
        unsigned long long add = t >= off ? part[t - off] : 0ull;
        __syncthreads();
        part[t] += add;
        __syncthreads();
    }
    if (t == 1023) total_s = part[1023];
    __syncthreads();
    const unsigned long long n = total_s;
    if (n == 0) {
        if (t == 0) {
            status[ST_HAS_NONZERO] = 0;
            status[ST_P95_BITS] = 0;
            status[ST_NZ_LO] = 0;
            status[ST_NZ_HI] = 0;
        }
        return;
    }
    // numpy: q = float32(95)/float32(100); vi = float32(n-1) * q   (all float32, round to nearest)
    const float q = __fdiv_rn(95.0f, 100.0f);
    const float nm1 = __ull2float_rn(n - 1);
    const float vi = __fmul_rn(nm1, q);
    float prev_f = floorf(vi);
    float next_f = __fadd_rn(prev_f, 1.0f);
    long long prev_i, next_i;
    if (vi >= nm1) {            // indexes_above_bounds -> index -1 = last element
        prev_i = next_i = (long long)n - 1;
    } else {
        prev_i = (long long)prev_f;
        next_i = (long long)next_f;
        if (next_i > (long long)n - 1) next_i = (long long)n - 1;   // cannot happen for q<1; guard
    }
    const float gamma = __fsub_rn(vi, prev_f);
    // value at sorted rank r = smallest bin with cumulative count > r
    const unsigned long long before = part[t] - mine;
    const long long ranks[2] = {prev_i, next_i};
    for (int k = 0; k < 2; ++k) {
        const unsigned long long r = (unsigned long long)ranks[k];
        if (r >= before && r < part[t]) {
            unsigned long long cum = before;
            for (int i = 0; i < 64; ++i) {
                const int b = b0 + i;
                if (b > pedestal) cum += ghist[b];
                if (cum > r) {
                    val_s[k] = b - pedestal;
                    break;
                }
            }
        }
    }
    __syncthreads();
    if (t == 0) {
        const float a = (float)val_s[0], b = (float)val_s[1];
        const float diff = __fsub_rn(b, a);
        float res = __fadd_rn(a, __fmul_rn(diff, gamma));
        if (gamma >= 0.5f) res = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, gamma)));
        status[ST_HAS_NONZERO] = 1;
        status[ST_P95_BITS] = __float_as_int(res);
        status[ST_NZ_LO] = (int32_t)(n & 0xffffffffull);
        status[ST_NZ_HI] = (int32_t)(n >> 32);
    }
}

int launch_percentile_finalize(tsp_handle* h, const uint32_t* d_hist, int pedestal, int32_t* d_status,
                               cudaStream_t s) {
    percentile_finalize_kernel<<<1, 1024, 0, s>>>(d_hist, pedestal, d_status);
    TSP_LAUNCH_CHECK(h);
    return TSP_OK;
}

// ---- K0: uint16 -> float32 with pedestal and p95 clip (SP:26-36) ------------------------------
__global__ void prepare_kernel(const uint16_t* __restrict__ in, float* __restrict__ out, size_t count,
                               int pedestal, const int32_t* __restrict__ status) {
    const bool clip = status[ST_HAS_NONZERO] != 0;
    const float p = __int_as_float(status[ST_P95_BITS]);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        int v = (int)in[i] - pedestal;
        float f = (float)(v > 0 ? v : 0);
        if (clip && f > p) f = p;
        out[i] = f;
    }
}

int launch_prepare(tsp_handle* h, const uint16_t* d_in, float* d_out, size_t count, int pedestal,
                   const int32_t* d_status, cudaStream_t s) {
    const int threads = 256;
    size_t blocks = (count + threads - 1) / threads;
    const size_t cap = (size_t)h->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    prepare_kernel<<<(int)blocks, threads, 0, s>>>(d_in, d_out, count, pedestal, d_status);
    TSP_LAUNCH_CHECK(h);
    return TSP_OK;
}

}  // namespace tsp
