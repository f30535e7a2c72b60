// Stage K1 - 95th percentile of the non-zero voxels (reference surface_projection.py:32-36), exact.
//
// numpy's 'linear' percentile needs the values at two adjacent sorted ranks; its index arithmetic
// is replayed in float32 (numpy/lib/_function_base_impl.py: virtual index (n-1)*q, floor, +1,
// gamma, _lerp - all float32 for float32 input; SURVEY trap T1).
//
// A full 65536-bin histogram costs one shared-memory atomic per voxel (~0.45 ms for 268 M voxels).
// Large volumes therefore take the order statistics in three steps, all exact:
//   1. sample_window_kernel: coarse histogram of a pseudo-random 1/stride subsample of 16-byte vectors; its
//      last CTA picks a value window [lo, lo+W) that contains the wanted ranks with overwhelming probability
//   2. window_count_kernel: one streaming pass that only COUNTS (voxels > pedestal, clamp sum -> voxels above
//      the window) and histograms the few voxels inside the window; its last CTA resolves the ranks inside
//      the window, or - if the window missed or was too wide - arms
//   3. hist_percentile_kernel, the full-histogram fallback (returns at once when not armed).
// Small volumes (stride 1) use the full histogram directly.  "Last CTA" = atomic ticket after a fence, so the
// whole percentile is three launches.
#include <type_traits>

#include "common.cuh"

namespace tsp {

constexpr int kHistThreads = 1024;
constexpr int kVecPerThread = 7;                                      // uint4 loads per thread per chunk
static_assert(kHistThreads * kVecPerThread * 8 + 16 < 65536, "16-bit counters must not overflow between flushes");
constexpr int kHistSmemBytes = kHistBins / 2 * 4;                     // 131072
constexpr int kWinBins = 4096;                                        // widest value window counted in one pass
constexpr size_t kSampleTarget = (size_t)1 << 22;                     // ~4 M sampled voxels

// ---- per-CTA privatised histogram: two 16-bit counters per 32-bit shared word --------------------
__device__ __forceinline__ void hist_add(uint32_t* sh, uint32_t v) {
    atomicAdd(&sh[v >> 1], (v & 1u) ? 0x10000u : 1u);
}

__device__ __forceinline__ void hist_add_word(uint32_t* sh, uint32_t w) {
    hist_add(sh, w & 0xffffu);
    hist_add(sh, w >> 16);
}

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// exact 65536-bin histogram of all `count` voxels into ghist (per-CTA privatised in shared memory)
__device__ void hist_full_body(const uint16_t* __restrict__ vol, size_t count, uint32_t* __restrict__ ghist,
                               uint32_t* sh) {
    for (int i = threadIdx.x; i < kHistBins / 2; i += kHistThreads) sh[i] = 0;
    __syncthreads();

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
    if (blockIdx.x == 0) {      // head/tail voxels when block 0 had no chunk to flush them with
        for (int i = threadIdx.x; i < kHistBins / 2; i += kHistThreads) {
            const uint32_t w = sh[i];
            if (w) {
                if (w & 0xffffu) atomicAdd(&ghist[i * 2], w & 0xffffu);
                if (w >> 16) atomicAdd(&ghist[i * 2 + 1], w >> 16);
            }
        }
    }
}

// ---- block-wide rank lookup in a table of counts --------------------------------------------------
// All threads of the block (a multiple of 32, at most 1024).  counts[0..L), L a multiple of 128 * warps; bins
// below `first_bin` are ignored.  Loads go to L2 (ld.cg): the table may have just been written by other CTAs.  Returns the
// total of the counted bins and, for each of the two ranks (0-based, after `base` earlier elements),
// the smallest bin whose cumulative count exceeds it (-1 when the rank is outside the table).
struct RankLookup {
    unsigned long long total;
    int bin[2];
};

__device__ RankLookup block_rank_lookup(const uint32_t* __restrict__ counts, int L, int first_bin,
                                        unsigned long long base, unsigned long long r0, unsigned long long r1) {
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long warp_off[33];
    __shared__ int found[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const int seg = L / nwarps;                               // bins per warp, multiple of 128
    const uint4* c4 = reinterpret_cast<const uint4*>(counts);
    if (threadIdx.x < 2) found[threadIdx.x] = -1;
    auto masked = [&](int b0, uint4 v) {
        if (b0 + 0 < first_bin) v.x = 0;
        if (b0 + 1 < first_bin) v.y = 0;
        if (b0 + 2 < first_bin) v.z = 0;
        if (b0 + 3 < first_bin) v.w = 0;
        return v;
    };
    unsigned long long mine = 0;
    for (int c = 0; c < seg / 128; ++c) {
        const int b0 = warp * seg + c * 128 + lane * 4;
        const uint4 v = masked(b0, __ldcg(c4 + b0 / 4));
        mine += (unsigned long long)v.x + v.y + v.z + v.w;
    }
    for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (lane == 0) warp_tot[warp] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long acc = base;
        for (int w = 0; w < nwarps; ++w) {
            warp_off[w] = acc;
            acc += warp_tot[w];
        }
        for (int w = nwarps; w <= 32; ++w) warp_off[w] = acc;
    }
    __syncthreads();
    const unsigned long long ranks[2] = {r0, r1};
    const unsigned long long wlo = warp_off[warp], whi = warp_off[warp + 1];
    if ((r0 >= wlo && r0 < whi) || (r1 >= wlo && r1 < whi)) {           // warp-uniform
        unsigned long long run = wlo;
        for (int c = 0; c < seg / 128; ++c) {
            const int b0 = warp * seg + c * 128 + lane * 4;
            const uint4 v = masked(b0, __ldcg(c4 + b0 / 4));
            const unsigned long long s4 = (unsigned long long)v.x + v.y + v.z + v.w;
            unsigned long long incl = s4;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            const unsigned long long excl = run + incl - s4;
            for (int k = 0; k < 2; ++k) {
                const unsigned long long r = ranks[k];
                if (r >= excl && r < excl + s4) {
                    unsigned long long cum = excl;
                    const uint32_t vs[4] = {v.x, v.y, v.z, v.w};
                    for (int i = 0; i < 4; ++i) {
                        cum += vs[i];
                        if (cum > r) {
                            found[k] = b0 + i;
                            break;
                        }
                    }
                }
            }
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
    RankLookup out;
    out.total = warp_off[32] - base;
    out.bin[0] = found[0];
    out.bin[1] = found[1];
    __syncthreads();
    return out;
}

// numpy 'linear' rank arithmetic in float32 for n values: ranks of the two neighbours and gamma
struct RankPair {
    unsigned long long prev, next;
    float gamma;
};

__device__ RankPair numpy_ranks(unsigned long long n, float q) {      // q = float32(percentile) / float32(100)
    const float nm1 = __ull2float_rn(n - 1);
    const float vi = __fmul_rn(nm1, q);
    const float prev_f = floorf(vi);
    const float next_f = __fadd_rn(prev_f, 1.0f);
    RankPair r;
    if (vi >= nm1) {                                   // indexes_above_bounds -> index -1 = last element
        r.prev = r.next = n - 1;
    } else {
        r.prev = (unsigned long long)prev_f;
        r.next = (unsigned long long)next_f;
        if (r.next > n - 1) r.next = n - 1;            // cannot happen for q < 1; guard
    }
    r.gamma = __fsub_rn(vi, prev_f);
    return r;
}

__device__ float numpy_lerp(float a, float b, float gamma) {
    const float diff = __fsub_rn(b, a);
    float res = __fadd_rn(a, __fmul_rn(diff, gamma));
    if (gamma >= 0.5f) res = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, gamma)));
    return res;
}

__device__ void write_result(int32_t* status, unsigned long long n, float p) {
    status[ST_HAS_NONZERO] = n > 0 ? 1 : 0;
    status[ST_P95_BITS] = __float_as_int(p);
    status[ST_NZ_LO] = (int32_t)(n & 0xffffffffull);
    status[ST_NZ_HI] = (int32_t)(n >> 32);
}

// true in exactly one CTA of the grid: the one that finishes last (its threads then see every other CTA's
// global writes and atomics).  All threads of the block must call it.
__device__ bool is_last_block(unsigned int* ticket) {
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    return last;
}

// ---- exact percentile from a full histogram (also the fallback) -----------------------------------
__device__ void percentile_finalize(const uint32_t* __restrict__ ghist, int pedestal, int32_t* __restrict__ status,
                                    float q) {
    // pass 1: total of the non-zero voxels, pass 2: the two ranks
    RankLookup first = block_rank_lookup(ghist, kHistBins, pedestal + 1, 0, ~0ull, ~0ull);
    const unsigned long long n = first.total;
    if (n == 0) {
        if (threadIdx.x == 0) write_result(status, 0, 0.f);
        return;
    }
    const RankPair rp = numpy_ranks(n, q);
    RankLookup look = block_rank_lookup(ghist, kHistBins, pedestal + 1, 0, rp.prev, rp.next);
    if (threadIdx.x == 0)
        write_result(status, n, numpy_lerp((float)(look.bin[0] - pedestal), (float)(look.bin[1] - pedestal), rp.gamma));
}

// SP:46 takes the percentile of ALL voxels of the second channel (zeros included): the voxels at or below the
// pedestal are `zeros` leading values 0 in the sorted order
__device__ void percentile_finalize_all(const uint32_t* __restrict__ ghist, int pedestal, unsigned long long count,
                                        int32_t* __restrict__ status, float q) {
    RankLookup first = block_rank_lookup(ghist, kHistBins, pedestal + 1, 0, ~0ull, ~0ull);
    const unsigned long long zeros = count - first.total;
    const RankPair rp = numpy_ranks(count, q);
    RankLookup look = block_rank_lookup(ghist, kHistBins, pedestal + 1, zeros, rp.prev, rp.next);
    if (threadIdx.x == 0) {
        const float v0 = look.bin[0] < 0 ? 0.f : (float)(look.bin[0] - pedestal);
        const float v1 = look.bin[1] < 0 ? 0.f : (float)(look.bin[1] - pedestal);
        write_result(status, count, numpy_lerp(v0, v1, rp.gamma));
    }
}

// full histogram of all voxels; the CTA that finishes last resolves the percentile.  `gate`: when non-null the
// kernel only runs if *gate != 0 (fallback armed by ST_NEED_FULL).
__global__ void __launch_bounds__(kHistThreads, 1)
hist_percentile_kernel(const uint16_t* __restrict__ vol, size_t count, uint32_t* __restrict__ ghist, int pedestal,
                       int32_t* __restrict__ status, unsigned int* __restrict__ ticket, const int32_t* __restrict__ gate,
                       int all_voxels, float q) {
    chain_release();
    chain_wait();
    if (gate && __ldcg(gate) == 0) return;
    extern __shared__ uint32_t sh[];
    hist_full_body(vol, count, ghist, sh);
    if (!is_last_block(ticket)) return;
    if (all_voxels) percentile_finalize_all(ghist, pedestal, (unsigned long long)count, status, q);
    else percentile_finalize(ghist, pedestal, status, q);
}

// ---- step 1: value window from a sample ------------------------------------------------------------
// One 128-byte line out of every `stride` (a power of two), at a hashed position inside its group - whole lines,
// because scattered 16-byte reads cost a DRAM transaction each; two of the eight voxels of every 16-byte vector
// (4 apart - neighbours are correlated) go into a coarse histogram of 8-value bins, non-zero voxels only.  The CTA that finishes last turns the sample ranks 0.95 (ns - 1) -/+ 9 sigma into the raw-value
// window [lo, lo + w) that window_count_kernel resolves exactly.
constexpr int kCoarseShift = 3;
constexpr int kCoarseBins = kHistBins >> kCoarseShift;      // 8192

// Run by the last CTA (kHistThreads threads): eight coarse bins per thread stay in registers - one round of loads
// gives the sample size, an exclusive scan and, once the rank interval is known, both of its ends.
static_assert(kCoarseBins == kHistThreads * 8, "window_select keeps eight coarse bins per thread");

__device__ void window_select(const uint32_t* __restrict__ shist, int pedestal, int32_t* __restrict__ status,
                              double q) {
    __shared__ unsigned long long warp_tot[kHistThreads / 32];
    __shared__ int found[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint4 b0 = __ldcg(reinterpret_cast<const uint4*>(shist) + 2 * threadIdx.x);
    const uint4 b1 = __ldcg(reinterpret_cast<const uint4*>(shist) + 2 * threadIdx.x + 1);
    const uint32_t bins[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    unsigned long long mine = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) mine += bins[i];
    unsigned long long incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) warp_tot[warp] = incl;
    if (threadIdx.x < 2) found[threadIdx.x] = -1;
    __syncthreads();
    unsigned long long before = 0, ns = 0;
    for (int w = 0; w < kHistThreads / 32; ++w) {
        if (w < warp) before += warp_tot[w];
        ns += warp_tot[w];
    }
    if (ns < 1024) {                  // sample too thin to trust: go straight to the full histogram
        if (threadIdx.x == 0) {
            status[ST_WIN_OK] = 0;
            status[ST_NEED_FULL] = 1;
        }
        return;
    }
    const double centre = q * (double)(ns - 1);
    const double margin = 4.0 + 9.0 * sqrt((double)ns * q * (1.0 - q));
    const double lo_r = centre - margin, hi_r = centre + margin + 1.0;
    const unsigned long long ranks[2] = {lo_r < 0.0 ? 0ull : (unsigned long long)lo_r,
                                         hi_r > (double)(ns - 1) ? ns - 1 : (unsigned long long)hi_r};
    const unsigned long long excl = before + incl - mine;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (ranks[r] >= excl && ranks[r] < excl + mine) {
            unsigned long long cum = excl;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                cum += bins[i];
                if (cum > ranks[r]) {
                    found[r] = 8 * threadIdx.x + i;
                    break;
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int lo = found[0] << kCoarseShift, hi = (found[1] << kCoarseShift) + (1 << kCoarseShift) - 1;
        if (lo < pedestal + 1) lo = pedestal + 1;
        if (hi > kHistBins - 1) hi = kHistBins - 1;
        int w = hi - lo + 1;
        const bool ok = found[0] >= 0 && found[1] >= 0 && w < kWinBins && w > 0;
        if (ok) w = (1 << (32 - __clz(w))) - 1;        // widened to 2^k - 1 values (clamp-and-sum count pass), k <= 12
        status[ST_WIN_LO] = lo;
        status[ST_WIN_N] = w;
        status[ST_WIN_OK] = ok ? 1 : 0;
        status[ST_NEED_FULL] = ok ? 0 : 1;
    }
}

__global__ void __launch_bounds__(kHistThreads, 1)
sample_window_kernel(const uint16_t* __restrict__ vol, size_t count, uint32_t stride, int pedestal,
                     uint32_t* __restrict__ shist, int32_t* __restrict__ status, unsigned int* __restrict__ ticket,
                     uint4* __restrict__ zero_ptr, size_t zero_vecs, double q) {
    chain_release();
    __shared__ uint32_t sh[kCoarseBins];
    for (int i = threadIdx.x; i < kCoarseBins; i += kHistThreads) sh[i] = 0;
    chain_wait();                // everything above overlaps the previous kernel's tail
    // side job: clear the accumulation volume of the score stage.  This kernel is latency bound (it reads 1/32 of the
    // volume), the stores ride along for free - as a memset node or inside the count pass they cost 4 us
    for (size_t i = (size_t)blockIdx.x * kHistThreads + threadIdx.x; i < zero_vecs; i += (size_t)gridDim.x * kHistThreads)
        zero_ptr[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const uintptr_t addr = reinterpret_cast<uintptr_t>(vol);
    size_t head = ((16 - (addr & 15)) & 15) / 2;
    if (head > count) head = count;
    const uint4* body = reinterpret_cast<const uint4*>(vol + head);
    const size_t nvec = ((count - head) / 8 / stride) & ~(size_t)7;      // sampled vectors: whole lines of 8
    const uint32_t ped = (uint32_t)pedestal;
    const size_t vec_per_chunk = (size_t)kHistThreads * kVecPerThread;
    const size_t nchunks = (nvec + vec_per_chunk - 1) / vec_per_chunk;
    for (size_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const size_t base = chunk * vec_per_chunk;
        uint2 v[kVecPerThread];
#pragma unroll
        for (int j = 0; j < kVecPerThread; ++j) {
            const size_t idx = base + (size_t)j * kHistThreads + threadIdx.x;
            // whole 128-byte lines (8 consecutive vectors, one per lane of an octet): one line out of every `stride`
            const size_t line = idx >> 3;
            const size_t src = ((line * stride + (hash32((uint32_t)line) & (stride - 1))) << 3) + (idx & 7);
            const uint4 q = idx < nvec ? __ldg(body + src) : make_uint4(0, 0, 0, 0);
            v[j] = make_uint2(q.x & 0xffffu, q.z & 0xffffu);
        }
#pragma unroll
        for (int j = 0; j < kVecPerThread; ++j) {
            if (v[j].x > ped) atomicAdd(&sh[v[j].x >> kCoarseShift], 1u);      // padding vectors hold zeros
            if (v[j].y > ped) atomicAdd(&sh[v[j].y >> kCoarseShift], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kCoarseBins; i += kHistThreads)
        if (sh[i]) atomicAdd(&shist[i], sh[i]);
    if (!is_last_block(ticket)) return;
    window_select(shist, pedestal, status, q);
}

// ---- step 2: streaming count pass -----------------------------------------------------------------
// One pass over the volume that only counts.  The window of step 1 is widened to W = 2^k - 1 values,
// [lo, lo + W), and every voxel is reduced to
//     c(v) = clamp(v, lo - 1, lo + W) - (lo - 1)  in [0, 2^k]:   0 below the window, 2^k above it, d + 1 at window bin d
// so ONE clamped value per voxel carries both facts the pass needs: a 16-byte vector touches the window iff some c
// has non-zero low k bits (one OR over the vector), and the sum R of all c gives the count above the window once
// the window's own share sum_d hist[d] (d + 1) is taken out:
//     above = (R - sum_d hist[d] (d + 1)) >> k,        below = count - in_window - above.
// Everything runs on packed uint16 pairs: VIMNMX.U16x2 max / min, a plain 32-bit subtract (no borrow: both halves
// are >= lo - 1), IDP.2A sums both halves into a 32-bit counter - 27 instructions per 16-byte vector.  The
// ~2 % of vectors that touch the window are parked in a shared queue and histogrammed after the stream.
// Measured on B200 (config 2, 537 MB): 80 us for the bare read stream of this very loop, +7 us for the counting.
constexpr int kCountThreads = 512;
constexpr int kCountUnroll = 4;
constexpr int kQueueCap = 1536;              // parked 16-byte vectors per CTA (24 KB)
static_assert(kWinBins == kCountThreads * 8, "window_finalize keeps eight window bins per thread");

// ---- step 3 (run by the last CTA of window_count_kernel) -----------------------------------------------
// Eight window bins per thread stay in registers: one round of loads gives the window total, its share of the
// clamp sum, an exclusive scan, and - once the counts below the window are known - both ranks.
__device__ void window_finalize(const uint32_t* __restrict__ gwin, const unsigned long long* __restrict__ gcounters,
                                unsigned long long count, int pedestal, int32_t* __restrict__ status, float q) {
    __shared__ unsigned long long warp_tot[kCountThreads / 32], warp_share[kCountThreads / 32];
    __shared__ int found[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long n = __ldcg(gcounters);      // voxels > pedestal
    const unsigned long long clamp_sum = __ldcg(gcounters + 1);
    const uint4 b0 = __ldcg(reinterpret_cast<const uint4*>(gwin) + 2 * threadIdx.x);
    const uint4 b1 = __ldcg(reinterpret_cast<const uint4*>(gwin) + 2 * threadIdx.x + 1);
    const uint32_t bins[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    unsigned long long mine = 0, share = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        mine += bins[i];
        share += (unsigned long long)bins[i] * (unsigned)(8 * threadIdx.x + i + 1);
    }
    unsigned long long incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    for (int o = 16; o; o >>= 1) share += __shfl_xor_sync(0xffffffffu, share, o);
    if (lane == 31) warp_tot[warp] = incl;
    if (lane == 0) warp_share[warp] = share;
    if (threadIdx.x < 2) found[threadIdx.x] = -1;
    __syncthreads();
    unsigned long long before = 0, in_window = 0, share_all = 0;
    for (int w = 0; w < kCountThreads / 32; ++w) {
        if (w < warp) before += warp_tot[w];
        in_window += warp_tot[w];
        share_all += warp_share[w];
    }
    if (n == 0) {
        if (threadIdx.x == 0) write_result(status, 0, 0.f);
        return;
    }
    const uint32_t wn = (uint32_t)st_load(status, ST_WIN_N);
    const int k = 32 - __clz(wn);                        // wn = 2^k - 1
    const unsigned long long zeros = count - n;          // voxels <= pedestal (they are all below the window)
    const bool sane = clamp_sum >= share_all && ((clamp_sum - share_all) & ((1ull << k) - 1)) == 0;
    const unsigned long long above = sane ? (clamp_sum - share_all) >> k : 0;
    const bool ok = sane && in_window + above + zeros <= count;
    const unsigned long long below_nz = ok ? count - in_window - above - zeros : 0;
    const RankPair rp = numpy_ranks(n, q);
    const unsigned long long ranks[2] = {rp.prev, rp.next};
    const unsigned long long excl = below_nz + before + incl - mine;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (ranks[r] >= excl && ranks[r] < excl + mine) {
            unsigned long long cum = excl;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                cum += bins[i];
                if (cum > ranks[r]) {
                    found[r] = 8 * threadIdx.x + i;
                    break;
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int lo = st_load(status, ST_WIN_LO);
        if (!ok || found[0] < 0 || found[1] < 0 || rp.prev < below_nz) {
            status[ST_NEED_FULL] = 1;               // the sample window missed: exact fallback
        } else {
            write_result(status, n, numpy_lerp((float)(lo + found[0] - pedestal), (float)(lo + found[1] - pedestal),
                                               rp.gamma));
        }
    }
}

__global__ void __launch_bounds__(kCountThreads, 2)
window_count_kernel(const uint16_t* __restrict__ vol, size_t count, int pedestal, int32_t* __restrict__ status,
                    uint32_t* __restrict__ gwin, unsigned long long* __restrict__ gcounters,
                    unsigned int* __restrict__ ticket, float q) {
    chain_release();
    __shared__ uint32_t win[kWinBins];
    __shared__ uint4 queue[kQueueCap];
    __shared__ uint32_t qtail;
    __shared__ unsigned long long blk[2];
    for (int i = threadIdx.x; i < kWinBins; i += kCountThreads) win[i] = 0;
    if (threadIdx.x < 2) blk[threadIdx.x] = 0;
    if (threadIdx.x == 0) qtail = 0;
    chain_wait();                // the shared-memory clears above overlap the sample kernel's last CTA
    if (st_load(status, ST_WIN_OK) == 0) return;
    const uint32_t lo = (uint32_t)st_load(status, ST_WIN_LO), wn = (uint32_t)st_load(status, ST_WIN_N);     // wn = 2^k - 1, lo >= 1
    const uint32_t ped = (uint32_t)pedestal;
    __syncthreads();

    unsigned long long nz_total = 0, clamp_total = 0;      // per-thread totals (scalar path + flushes)
    auto one = [&](uint32_t v) {
        if (v > ped) ++nz_total;
        if (v >= lo) {
            const uint32_t d = v - lo;
            clamp_total += d < wn ? d + 1 : wn + 1;
            if (d < wn) atomicAdd(&win[d], 1u);
        }
    };
    const uintptr_t addr = reinterpret_cast<uintptr_t>(vol);
    size_t head = ((16 - (addr & 15)) & 15) / 2;
    if (head > count) head = count;
    const size_t nvec = (count - head) / 8;
    const size_t tail0 = head + nvec * 8;
    const uint4* body = reinterpret_cast<const uint4*>(vol + head);
    if (blockIdx.x == 0) {
        for (size_t i = threadIdx.x; i < head; i += kCountThreads) one(vol[i]);
        for (size_t i = tail0 + threadIdx.x; i < count; i += kCountThreads) one(vol[i]);
    }
    auto park = [&](const uint4& vec) {      // rare: the vector goes to the CTA queue, histogrammed after the stream
        const uint32_t slot = atomicAdd(&qtail, 1u);
        if (slot < kQueueCap) {
            queue[slot] = vec;
        } else {
            const uint32_t ws[4] = {vec.x, vec.y, vec.z, vec.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t d0 = (ws[q] & 0xffffu) - lo, d1 = (ws[q] >> 16) - lo;
                if (d0 < wn) atomicAdd(&win[d0], 1u);
                if (d1 < wn) atomicAdd(&win[d1], 1u);
            }
        }
    };
    // packed clamp bounds; the top bound saturates at 65535 (then nothing lies above the window)
    const uint32_t lo1_2 = (lo - 1) * 0x00010001u;
    const uint32_t hi_2 = min(lo + wn, 65535u) * 0x00010001u;
    const uint32_t mask_2 = wn * 0x00010001u;
    const uint32_t ped1_2 = (ped + 1) * 0x00010001u, ped_2 = ped * 0x00010001u;
    uint32_t acc_c = 0, acc_nz = 0, rounds = 0;
    auto flush = [&]() {
        clamp_total += acc_c;
        nz_total += acc_nz;
        acc_c = acc_nz = 0;
    };
    // chunk c = vectors [c * T * U, (c + 1) * T * U): thread t takes c * T * U + j * T + t, j < U (no bounds checks)
    constexpr size_t kChunkVec = (size_t)kCountThreads * kCountUnroll;
    const size_t nfull = nvec / kChunkVec;
    auto stream = [&](auto zero_pedestal_tag) {
        constexpr bool ZERO_PED = decltype(zero_pedestal_tag)::value;       // [v > 0] is min(v, 1)
        for (size_t chunk = blockIdx.x; chunk < nfull; chunk += gridDim.x) {
            const uint4* src = body + chunk * kChunkVec + threadIdx.x;
            uint4 v[kCountUnroll];
#pragma unroll
            for (int j = 0; j < kCountUnroll; ++j) v[j] = __ldg(src + j * kCountThreads);
#pragma unroll
            for (int j = 0; j < kCountUnroll; ++j) {
                const uint32_t ws[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
                uint32_t c[4], f[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    c[q] = __vminu2(__vmaxu2(ws[q], lo1_2), hi_2) - lo1_2;
                    f[q] = ZERO_PED ? __vminu2(ws[q], 0x00010001u) : __vminu2(__vmaxu2(ws[q], ped_2), ped1_2) - ped_2;
                }
                acc_c = __dp2a_lo((c[0] + c[1]) + (c[2] + c[3]), 0x0101u, acc_c);      // halves <= 4 * 4096: no carry
                acc_nz = __dp2a_lo((f[0] + f[1]) + (f[2] + f[3]), 0x0101u, acc_nz);
                if (((c[0] | c[1]) | (c[2] | c[3])) & mask_2) park(v[j]);
            }
            if (++rounds == 8192u) {           // 32-bit counters: at most 8 * kCountUnroll * 4096 per round
                flush();
                rounds = 0;
            }
        }
    };
    if (ped == 0) stream(std::true_type{});
    else stream(std::false_type{});
    if (blockIdx.x == 0) {                 // the vectors after the last full chunk, voxel by voxel
        const uint16_t* rest = reinterpret_cast<const uint16_t*>(body + nfull * kChunkVec);
        const size_t nrest = (nvec - nfull * kChunkVec) * 8;
        for (size_t i = threadIdx.x; i < nrest; i += kCountThreads) one(rest[i]);
    }
    flush();
    __syncthreads();
    {   // drain the queue: one parked vector per thread per round, every lane busy
        const uint32_t nq = min(qtail, (uint32_t)kQueueCap);
        for (uint32_t e = threadIdx.x; e < nq; e += kCountThreads) {
            const uint4 qv = queue[e];
            const uint32_t ws[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint32_t val = (q & 1) ? (ws[q >> 1] >> 16) : (ws[q >> 1] & 0xffffu);
                const uint32_t d = val - lo;
                if (d < wn) atomicAdd(&win[d], 1u);
            }
        }
    }
    unsigned long long nz = nz_total, cs = clamp_total;
    for (int o = 16; o; o >>= 1) {
        nz += __shfl_xor_sync(0xffffffffu, nz, o);
        cs += __shfl_xor_sync(0xffffffffu, cs, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&blk[0], nz);
        atomicAdd(&blk[1], cs);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(&gcounters[0], blk[0]);
        atomicAdd(&gcounters[1], blk[1]);
    }
    for (int i = threadIdx.x; i <= (int)wn && i < kWinBins; i += kCountThreads)
        if (win[i]) atomicAdd(&gwin[i], win[i]);
    if (!is_last_block(ticket)) return;
    window_finalize(gwin, gcounters, (unsigned long long)count, pedestal, status, q);
}

// ---- launchers ------------------------------------------------------------------------------------
size_t percentile_scratch_bytes() {
    // [full histogram 256 KB][sample histogram 32 KB][window 16 KB][counters + tickets 64 B]
    return (kHistBins + kCoarseBins + kWinBins) * sizeof(uint32_t) + 64;
}

static int ensure_hist_attr(tsp_handle* h) {
    if (!h->hist_attr) {
        TSP_CUDA(cudaFuncSetAttribute(hist_percentile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHistSmemBytes));
        h->hist_attr = true;
    }
    return TSP_OK;
}

static int hist_grid(tsp_handle* h, size_t count, uint32_t stride) {
    const size_t per_chunk = (size_t)kHistThreads * kVecPerThread * 8 * stride;
    const size_t nchunks = (count + per_chunk - 1) / per_chunk;
    return (int)(nchunks < (size_t)h->sm_count ? (nchunks ? nchunks : 1) : (size_t)h->sm_count);
}

// d_scratch: percentile_scratch_bytes() bytes.  Writes ST_HAS_NONZERO / ST_P95_BITS / ST_NZ_* into d_status.
// Three launches: sample + window, exact count + resolve, gated full-histogram fallback.
// Also clears the status block (kStatusWords words) and - for the next stage - zero_bytes at zero_ptr (16-byte
// aligned, may be null).
int launch_percentile(tsp_handle* h, const uint16_t* d_vol, size_t count, int pedestal, int32_t* d_status,
                      void* d_scratch, cudaStream_t s, void* zero_ptr, size_t zero_bytes, float q, double q64) {
    int rc = ensure_hist_attr(h);
    if (rc) return rc;
    uint32_t* hist_full = (uint32_t*)d_scratch;
    uint32_t* hist_sample = hist_full + kHistBins;
    uint32_t* win = hist_sample + kCoarseBins;
    unsigned long long* counters = (unsigned long long*)(win + kWinBins);
    unsigned int* tickets = (unsigned int*)(counters + 2);
    if ((char*)d_scratch - (char*)d_status == (ptrdiff_t)align_up(kStatusWords * sizeof(int32_t), 256)) {
        TSP_CUDA(cudaMemsetAsync(d_status, 0, align_up(kStatusWords * sizeof(int32_t), 256) + percentile_scratch_bytes(), s));
    } else {
        TSP_CUDA(cudaMemsetAsync(d_status, 0, kStatusWords * sizeof(int32_t), s));
        TSP_CUDA(cudaMemsetAsync(d_scratch, 0, percentile_scratch_bytes(), s));
    }
    const bool zero_vec_ok = zero_ptr && (reinterpret_cast<uintptr_t>(zero_ptr) & 15) == 0 && zero_bytes % 16 == 0;
    // sampling stride: ~4 M sampled lines' worth of voxels for large volumes, never less than ~1 M (a 16 MiB stack is
    // sampled 1 in 8: the full-histogram kernel needs 35 us there - its 148 privatised tables flush 65536 bins each -
    // the sampled window + count pass 20 us); below 2 M voxels the full histogram is the cheaper way
    size_t target = count / 16;
    if (target < ((size_t)1 << 19)) target = (size_t)1 << 19;
    if (target > kSampleTarget) target = kSampleTarget;
    uint32_t stride = 1;
    while ((count / stride) > 2 * target && stride < 1024) stride *= 2;
    if (stride == 1 || !zero_vec_ok) {
        if (zero_ptr && zero_bytes) TSP_CUDA(cudaMemsetAsync(zero_ptr, 0, zero_bytes, s));
        zero_ptr = nullptr;
        zero_bytes = 0;
    }
    if (stride == 1) {
        hist_percentile_kernel<<<hist_grid(h, count, 1), kHistThreads, kHistSmemBytes, s>>>(
            d_vol, count, hist_full, pedestal, d_status, tickets, nullptr, 0, q);
        TSP_LAUNCH_CHECK(h);
        return TSP_OK;
    }
    TSP_CUDA(launch_after_copy(sample_window_kernel, hist_grid(h, count, stride), kHistThreads, 0, s, d_vol, count, stride,
                            pedestal, hist_sample, d_status, tickets + 1, (uint4*)zero_ptr, zero_bytes / 16, q64));
    TSP_LAUNCH_CHECK(h);
    prof_mark(h, s, STG_PCT_SAMPLE);
    TSP_CUDA(launch_chained(window_count_kernel, h->sm_count * 2, kCountThreads, 0, s, d_vol, count, pedestal, d_status,
                            win, counters, tickets + 2, q));
    TSP_LAUNCH_CHECK(h);
    prof_mark(h, s, STG_PCT_COUNT);      // what follows (the gated fallback) is booked on the caller's STG_PERCENTILE
    // exact fallback, armed by ST_NEED_FULL (returns at once otherwise).  It is almost never armed (a sample of fewer
    // than 1024 non-zero voxels, a window wider than 4096 values), but its CTAs - 1024 threads, 128 KB of shared
    // memory - each claim a whole SM while they are scheduled: a small grid keeps the gate cheap for the frames
    // running next to this one (the armed case then takes 16 instead of 148 SMs for its one pass over the stack)
    int fgrid = hist_grid(h, count, 1);
    if (fgrid > 16) fgrid = 16;
    TSP_CUDA(launch_chained(hist_percentile_kernel, fgrid, kHistThreads, kHistSmemBytes, s, d_vol, count,
                            hist_full, pedestal, d_status, tickets, d_status + ST_NEED_FULL, 0, q));
    TSP_LAUNCH_CHECK(h);
    return TSP_OK;
}

// SP:46: np.percentile(channel, 95) over every voxel (after the pedestal), exact, by the full histogram.
// d_status: a status block of its own (ST_P95_BITS / ST_HAS_NONZERO are what launch_prepare reads).
int launch_percentile_all(tsp_handle* h, const uint16_t* d_vol, size_t count, int pedestal, int32_t* d_status,
                          void* d_scratch, cudaStream_t s, float q) {
    int rc = ensure_hist_attr(h);
    if (rc) return rc;
    uint32_t* hist_full = (uint32_t*)d_scratch;
    unsigned int* tickets = (unsigned int*)((unsigned long long*)(hist_full + kHistBins + kCoarseBins + kWinBins) + 2);
    TSP_CUDA(cudaMemsetAsync(d_status, 0, kStatusWords * sizeof(int32_t), s));
    TSP_CUDA(cudaMemsetAsync(d_scratch, 0, percentile_scratch_bytes(), s));
    hist_percentile_kernel<<<hist_grid(h, count, 1), kHistThreads, kHistSmemBytes, s>>>(
        d_vol, count, hist_full, pedestal, d_status, tickets, nullptr, 1, q);
    TSP_LAUNCH_CHECK(h);
    return TSP_OK;
}

// ---- K0: uint16 -> float32 with pedestal and p95 clip (SP:26-36) ------------------------------
__global__ void prepare_kernel(const uint16_t* __restrict__ in, float* __restrict__ out, size_t count,
                               int pedestal, const int32_t* __restrict__ status) {
    const bool clip = st_load(status, ST_HAS_NONZERO) != 0;
    const float p = __int_as_float(st_load(status, ST_P95_BITS));
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
