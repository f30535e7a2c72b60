// Stages K4-K7: argmax over z (SP:61), band mask (SP:62-71) and weighted max projection (SP:72-81).
//
// The reference scatters a one-hot (Z, Y*X) volume at the height map, blurs it with sigma=(1,2,2)
// and takes max_z(image * mask).  The band kernels never materialise either volume: for a
// 32x64 pixel tile they walk only the planes near the tile's heights, builds the XY-blurred
// one-hot plane A[z''] on the fly from the height-map tile (x pass into shared memory, y pass into
// registers), keeps a 9-plane window of A per pixel and applies the z taps as the exact
// edge-replicating 9-band matrix (SURVEY trap T9), multiplying the raw uint16 voxels as they
// stream by.  Only planes inside the band are read from HBM.
#include <type_traits>

#include "common.cuh"

namespace tsp {

__constant__ float c_w2[17];      // sigma = 2 taps (SP:70, axes y and x)

// ---- K4 ----------------------------------------------------------------------------------------
__global__ void argmax_z_kernel(const float* __restrict__ score, int32_t* __restrict__ zmap, int Z,
                                size_t plane, int z_offset, int32_t* __restrict__ status) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    int zmin = INT32_MAX, zmax = INT32_MIN;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < plane; p += stride) {
        float best = score[p];
        int bz = 0;
        for (int z = 1; z < Z; ++z) {
            const float v = score[(size_t)z * plane + p];
            if (v > best) {          // strict: first maximum wins, like np.argmax
                best = v;
                bz = z;
            }
        }
        zmap[p] = bz + z_offset;
        zmin = min(zmin, bz + z_offset);
        zmax = max(zmax, bz + z_offset);
    }
    if (status) {
        for (int o = 16; o; o >>= 1) {
            zmin = min(zmin, __shfl_xor_sync(0xffffffffu, zmin, o));
            zmax = max(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
        }
        if ((threadIdx.x & 31) == 0 && zmax >= zmin) {
            atomicMax(&status[ST_ZMAX], zmax);
            atomicMax(&status[ST_ZMIN_INV], INT32_MAX - zmin);
        }
    }
}

int launch_argmax(tsp_handle* h, const float* d_score, int32_t* d_zmap, int Z, int Y, int X,
                  int z_offset, int32_t* d_status, cudaStream_t s) {
    const size_t plane = (size_t)Y * X;
    const int threads = 256;
    size_t blocks = (plane + threads - 1) / threads;
    if (blocks > (size_t)h->sm_count * 32) blocks = (size_t)h->sm_count * 32;
    argmax_z_kernel<<<(int)blocks, threads, 0, s>>>(d_score, d_zmap, Z, plane, z_offset, d_status);
    TSP_LAUNCH_CHECK(h);
    return TSP_OK;
}

// ---- height-map range + IndexError condition (SP:62, SP:68-69) -------------------------------
__global__ void zmap_range_kernel(const int32_t* __restrict__ zmap, size_t n, int32_t* __restrict__ status) {
    int lo = INT32_MAX, hi = INT32_MIN;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int v = zmap[i];
        lo = min(lo, v);
        hi = max(hi, v);
    }
    for (int o = 16; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&status[ST_ZMIN], lo);
        atomicMax(&status[ST_ZMAX], hi);
    }
}

__global__ void zmap_range_init_kernel(int32_t* status) {
    status[ST_ZMIN] = INT32_MAX;
    status[ST_ZMAX] = INT32_MIN;
}

// the reference indexes the cropped stack with chosen_z (and clip(chosen_z+shift, 0, Z), upper bound
// inclusive): any index >= Z raises IndexError; negative indices cannot occur.
__global__ void band_check_kernel(int32_t* status, int Z, int shift, int decode) {
    if (decode) status[ST_ZMIN] = INT32_MAX - status[ST_ZMIN_INV];     // argmax kernels keep max(INT_MAX - z)
    const int hi = st_load(status, ST_ZMAX);
    int err = hi >= Z;
    if (shift != 0) {
        int hs = hi + shift;
        hs = hs < 0 ? 0 : (hs > Z ? Z : hs);
        err |= hs >= Z;
    }
    if (st_load(status, ST_ZMIN) < 0) err = 1;
    status[ST_BAND_ERR] = err;
}

// ---- K5-K7 fused -------------------------------------------------------------------------------
constexpr int kBandTY = 32, kBandTX = 64, kBandHalo = 8;
constexpr int kBandThreads = 256;           // thread = 8 consecutive pixels of one row
constexpr int kBandPix = 8;
constexpr int kBandMaxCh = 2;               // channels per CTA
constexpr int kBandMaxPlanes = 4096;

__constant__ float c_w1[9];       // sigma = 1 taps (z axis, interior planes)

struct BandArgs {
    const uint16_t* stack;      // (C, Zfull, Y, X)
    size_t channel_stride;      // Zfull*Y*X
    size_t z0_offset;           // min_z*Y*X : first plane of the cropped stack
    const int32_t* zmap;        // (Y, X) indices into the cropped stack (before shift)
    float* proj;                // (C, Y, X)
    const float* wz;            // (Z, 9) edge-replicating z taps, wz[z*9+i] multiplies A[z-4+i]
    const int32_t* status;
    int Z, Y, X;
    int shift;
    int pedestal;
    int nch;
    int vec;                    // rows are 16-byte aligned: uint4 loads
    int zvec;                   // height-map rows are 16-byte aligned: int4 loads
    const float* lut;           // x-blur lookup tables of a binary row: [512] taps 0..8, [256] taps 9..16
    int* worklist;              // linear tile ids left to the deep-range kernel, status[ST_WORK_COUNT] of them (may be null)
    int use_list;               // band_project3_kernel: take the tiles from the worklist instead of blockIdx
    int tiles_x;                // tiles per image row (decodes worklist entries)
    int fused_check;            // 1: the IndexError rule is evaluated here from the argmax stage's range (no check kernel)
    int err_shift;              // atoh_shift of the frame (the rule looks at it even in the un-shifted pass)
    int ch[16];
};

// the reference indexes the cropped stack with chosen_z (and clip(chosen_z+shift, 0, Z), upper bound inclusive):
// any index >= Z raises IndexError; negative indices cannot occur.  Every CTA evaluates the rule from the range the
// argmax stage left in the status block (ST_ZMAX, ST_ZMIN_INV = max(INT_MAX - z)); the first CTA records the verdict.
__device__ __forceinline__ bool band_index_error(const BandArgs& a) {
    if (!a.fused_check) return st_load(a.status, ST_BAND_ERR) != 0;
    const int hi = st_load(a.status, ST_ZMAX), lo = INT32_MAX - st_load(a.status, ST_ZMIN_INV);
    int err = hi >= a.Z;
    if (a.err_shift != 0) err |= min(max(hi + a.err_shift, 0), a.Z) >= a.Z;
    if (lo < 0) err = 1;
    if (threadIdx.x == 0 && (blockIdx.x | blockIdx.y | blockIdx.z) == 0) {
        int32_t* st = const_cast<int32_t*>(a.status);
        st[ST_ZMIN] = lo;
        st[ST_BAND_ERR] = err;
    }
    return err != 0;
}

__device__ __forceinline__ uint4 band_load8(const uint16_t* __restrict__ src, int x, int X, bool vec) {
    if (vec && x + 7 < X) return __ldg(reinterpret_cast<const uint4*>(src));
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < 8; ++c)
        if (x + c < X) w[c >> 1] |= (uint32_t)__ldg(src + c) << (16 * (c & 1));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// ---- K5-K7 fused, second generation -------------------------------------------------------------
// Register-prefetch kernel (no TMA; any alignment): the per-plane work is cut to what the data need:
//   * the indicator [cz == t] of a tile row is an 80-bit mask (three warp ballots); the 17-tap x blur of a
//     binary line is a table lookup: window bits 0..8 and 9..16 index two shared tables of partial tap sums
//     (2 LDS + 1 FADD per pixel instead of 17 FMA and 17 compares); all-zero / all-one windows are free;
//   * per-plane shared buffers are double buffered: ONE block barrier per present plane;
//   * a 48-bit column mask per pixel group says which rows of the x-blurred plane are non-zero: warps whose
//     17-row windows see nothing skip the y pass of that plane altogether;
//   * y pass, z scatter and the weighted max use packed FFMA2 / FMUL2 (two pixels per instruction).
constexpr int kB2Threads = 256;
constexpr int kB2CW = kBandTX + 2 * kBandHalo;      // 80
constexpr int kB2CH = kBandTY + 2 * kBandHalo;      // 48

__constant__ float2 c_w2p[17];    // (w, w) pairs of the sigma = 2 taps
__constant__ float2 c_w1p[9];     // (w, w) pairs of the sigma = 1 taps

__device__ __forceinline__ float band_u16f_lo(uint32_t w, uint32_t magic) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, 0x7410;" : "=r"(r) : "r"(w), "r"(magic));
    return __uint_as_float(r);
}
__device__ __forceinline__ float band_u16f_hi(uint32_t w, uint32_t magic) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, 0x7432;" : "=r"(r) : "r"(w), "r"(magic));
    return __uint_as_float(r);
}

template <bool AIRY, bool TWO>
__global__ void __launch_bounds__(kB2Threads, 2) band_project2_kernel(const BandArgs a) {
    __shared__ __align__(16) int cz_s[kB2CH][kB2CW];
    __shared__ __align__(16) uint32_t rowmask[2][kB2CH][4];
    __shared__ __align__(16) float r_s[2][kB2CH][kBandTX];       // row layout: [half][group][4]
    __shared__ unsigned long long colbits[3][8];
    __shared__ float lut_lo[512], lut_hi[256];
    __shared__ uint32_t present[kBandMaxPlanes / 32];
    __shared__ int zlo_s, zhi_s;

    chain_release();
    chain_wait();
    if (band_index_error(a)) return;            // the reference raises before projecting

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * kBandTX, y0 = blockIdx.y * kBandTY;
    if (tid == 0) {
        zlo_s = INT32_MAX;
        zhi_s = INT32_MIN;
    }
    for (int i = tid; i < (a.Z >> 5) + 1 && i < kBandMaxPlanes / 32; i += kB2Threads) present[i] = 0;
    for (int i = tid; i < 512; i += kB2Threads) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) s += ((i >> k) & 1) ? c_w2[k] : 0.f;
        lut_lo[i] = s;
    }
    {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += ((tid >> k) & 1) ? c_w2[9 + k] : 0.f;
        lut_hi[tid] = s;
    }
    if (tid < 24) colbits[tid >> 3][tid & 7] = 0ull;
    __syncthreads();
    {
        constexpr int kPer = (kB2CH * kB2CW + kB2Threads - 1) / kB2Threads;
        int vals[kPer];
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int i = tid + k * kB2Threads;
            const int yy = min(max(y0 - kBandHalo + i / kB2CW, 0), a.Y - 1);
            const int xx = min(max(x0 - kBandHalo + i % kB2CW, 0), a.X - 1);
            vals[k] = i < kB2CH * kB2CW ? __ldg(a.zmap + (size_t)yy * a.X + xx) : 0;
        }
        int lo = INT32_MAX, hi = INT32_MIN;
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int i = tid + k * kB2Threads;
            if (i < kB2CH * kB2CW) {
                int v = vals[k];
                if (a.shift != 0) v = min(max(v + a.shift, 0), a.Z);
                cz_s[i / kB2CW][i % kB2CW] = v;
                lo = min(lo, v);
                hi = max(hi, v);
                atomicOr(&present[v >> 5], 1u << (v & 31));
            }
        }
        for (int o = 16; o; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) {
            atomicMin(&zlo_s, lo);
            atomicMax(&zhi_s, hi);
        }
    }
    __syncthreads();
    const int zlo = zlo_s, zhi = zhi_s;
    const float one_val = lut_lo[511] + lut_hi[255];

    const int g = tid & 7, row = tid >> 3;                     // 8 pixels x0+8g .. +7 of row y0+row
    const int x = x0 + g * kBandPix, y = y0 + row;
    const bool inside = y < a.Y && x < a.X;
    const int c0 = blockIdx.z * kBandMaxCh;
    const size_t plane = (size_t)a.Y * a.X;
    const uint16_t* src0 = a.stack + (size_t)a.ch[c0] * a.channel_stride + a.z0_offset + (size_t)y * a.X + x;
    const uint16_t* src1 = a.stack + (size_t)a.ch[c0 + (TWO ? 1 : 0)] * a.channel_stride + a.z0_offset +
                           (size_t)y * a.X + x;
    const float nb = -((float)a.pedestal + 8388608.0f);
    const float2 nbias = make_float2(nb, nb);
    uint32_t magic = 0x4B000000u;
    asm volatile("" : "+r"(magic));

    // ballots of plane t into rowmask[buf] (+ the zeroed column mask of that plane)
    auto build_masks = [&](int t, int buf, int cb) {
#pragma unroll
        for (int rr = 0; rr < kB2CH / 8; ++rr) {
            const int r = warp + 8 * rr;
            uint32_t mine = 0;
#pragma unroll
            for (int seg = 0; seg < 3; ++seg) {
                const int col = 32 * seg + lane;
                const int v = col < kB2CW ? cz_s[r][col] : -1;
                const uint32_t b = __ballot_sync(0xffffffffu, v == t);
                if (lane == seg) mine = b;
            }
            if (lane < 4) rowmask[buf][r][lane] = mine;          // word 3 = 0
        }
        if (tid < 8) colbits[cb][tid] = 0ull;
    };
    auto next_present = [&](int t) {                             // smallest present plane > t, or INT_MAX
        for (int q = t + 1; q <= zhi; ++q)
            if ((present[q >> 5] >> (q & 31)) & 1u) return q;
        return INT32_MAX;
    };
    // x pass of the plane whose masks are in rowmask[buf]: r_s[buf] and colbits[cb]
    auto x_pass = [&](int buf, int cb) {
#pragma unroll
        for (int round = 0; round < 2; ++round) {
            const int r = round * 32 + row;
            if (round == 1 && r >= kB2CH) break;                 // warp-uniform (warps 0..3 take the second round)
            const uint4 m = *reinterpret_cast<const uint4*>(&rowmask[buf][r][0]);
            const uint32_t wlo = g < 4 ? m.x : m.y, whi = g < 4 ? m.y : m.z;
            const uint32_t W = __funnelshift_r(wlo, whi, 8 * (g & 3)) & 0xFFFFFFu;
            float o[8];
            if (W == 0u) {
#pragma unroll
                for (int p = 0; p < 8; ++p) o[p] = 0.f;
            } else if (W == 0xFFFFFFu) {
#pragma unroll
                for (int p = 0; p < 8; ++p) o[p] = one_val;
            } else {
#pragma unroll
                for (int p = 0; p < 8; ++p) o[p] = lut_lo[(W >> p) & 511u] + lut_hi[(W >> (p + 9)) & 255u];
            }
            float4* dst = reinterpret_cast<float4*>(&r_s[buf][r][0]);
            dst[g] = make_float4(o[0], o[1], o[2], o[3]);
            dst[8 + g] = make_float4(o[4], o[5], o[6], o[7]);
            const uint32_t B = __ballot_sync(0xffffffffu, W != 0u);          // bit 8*rr + g, rows 4*warp + rr
            if (lane < 8) {
                const uint32_t nib = ((B >> lane) & 1u) | (((B >> (lane + 8)) & 1u) << 1) |
                                     (((B >> (lane + 16)) & 1u) << 2) | (((B >> (lane + 24)) & 1u) << 3);
                if (nib) atomicOr(&colbits[cb][lane], (unsigned long long)nib << (round * 32 + 4 * warp));
            }
        }
    };

    float2 win[4][11];            // pending masks of planes t-4 .. t+4(+2), two pixels per entry
    float2 best0[4], best1[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int i = 0; i < 11; ++i) win[q][i] = make_float2(0.f, 0.f);
        best0[q] = best1[q] = make_float2(0.f, 0.f);
    }
    int live = 0;

    uint4 nxt0 = make_uint4(0, 0, 0, 0), nxt1 = nxt0;
    {
        const int z = zlo - 4;
        if (inside && z >= 0 && z < a.Z) {
            nxt0 = band_load8(src0 + (size_t)z * plane, x, a.X, a.vec != 0);
            if (TWO) nxt1 = band_load8(src1 + (size_t)z * plane, x, a.X, a.vec != 0);
        }
    }

    int buf = 0, cb = 0;
    build_masks(zlo, 0, 0);                      // zlo is always present
    __syncthreads();

    for (int tb = zlo; tb <= zhi + 8; tb += 3) {
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int t = tb + u;
            if (t > zhi + 8) break;                                    // block-uniform
            const uint4 cur0 = nxt0, cur1 = nxt1;
            {   // prefetch the voxels of the next output plane while this one is processed
                const int zn = t - 3;
                if (inside && zn >= 0 && zn < a.Z && t + 1 <= zhi + 8) {
                    nxt0 = band_load8(src0 + (size_t)zn * plane, x, a.X, a.vec != 0);
                    if (TWO) nxt1 = band_load8(src1 + (size_t)zn * plane, x, a.X, a.vec != 0);
                }
            }
            const bool have = t <= zhi && ((present[t >> 5] >> (t & 31)) & 1u);
            if (have) {                                                // block-uniform
                x_pass(buf, cb);
                const int tn = next_present(t);
                const int cbn = cb == 2 ? 0 : cb + 1;
                if (tn != INT32_MAX) build_masks(tn, buf ^ 1, cbn);
                __syncthreads();
                // y pass: a_new = sum_dy w2[dy] * r_s[row + dy][8g ..], skipped by warps that see only zeros
                const bool mine = ((colbits[cb][g] >> row) & 0x1FFFFull) != 0ull;
                if (__any_sync(0xffffffffu, mine)) {
                    float2 an[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) an[q] = make_float2(0.f, 0.f);
#pragma unroll
                    for (int dy = 0; dy < 17; ++dy) {
                        const float4* src = reinterpret_cast<const float4*>(&r_s[buf][row + dy][0]);
                        const float4 lo4 = src[g], hi4 = src[8 + g];
                        const float2 w = c_w2p[dy];
                        an[0] = __ffma2_rn(make_float2(lo4.x, lo4.y), w, an[0]);
                        an[1] = __ffma2_rn(make_float2(lo4.z, lo4.w), w, an[1]);
                        an[2] = __ffma2_rn(make_float2(hi4.x, hi4.y), w, an[2]);
                        an[3] = __ffma2_rn(make_float2(hi4.z, hi4.w), w, an[3]);
                    }
                    if (mine) {
                        live = 9;
                        // plane t feeds the masks of z = t-4+k (window slot u+k) with the (z, t) entry of the
                        // edge-replicating z matrix: interior rows are the plain sigma=1 taps
                        const bool edge = t - 4 < 4 || t + 4 > a.Z - 5;
#pragma unroll
                        for (int k = 0; k < 9; ++k) {
                            const int z = t - 4 + k;
                            float2 w = c_w1p[8 - k];
                            if (edge) {
                                const float we = (z >= 0 && z < a.Z) ? __ldg(a.wz + z * 9 + (8 - k)) : 0.f;
                                w = make_float2(we, we);
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) win[q][u + k] = __ffma2_rn(w, an[q], win[q][u + k]);
                        }
                    }
                }
                buf ^= 1;
                cb = cbn;
            }
            const int z = t - 4;
            if (live > 0 && inside && z >= 0 && z < a.Z) {
                const uint32_t w0[4] = {cur0.x, cur0.y, cur0.z, cur0.w};
                const uint32_t w1[4] = {cur1.x, cur1.y, cur1.z, cur1.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 m = win[q][u];
                    float2 f = __fadd2_rn(make_float2(band_u16f_lo(w0[q], magic), band_u16f_hi(w0[q], magic)), nbias);
                    if (AIRY) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); }
                    const float2 pr = __fmul2_rn(f, m);
                    best0[q].x = fmaxf(best0[q].x, pr.x);
                    best0[q].y = fmaxf(best0[q].y, pr.y);
                    if (TWO) {
                        float2 f1 = __fadd2_rn(make_float2(band_u16f_lo(w1[q], magic), band_u16f_hi(w1[q], magic)), nbias);
                        if (AIRY) { f1.x = fmaxf(f1.x, 0.f); f1.y = fmaxf(f1.y, 0.f); }
                        const float2 pr1 = __fmul2_rn(f1, m);
                        best1[q].x = fmaxf(best1[q].x, pr1.x);
                        best1[q].y = fmaxf(best1[q].y, pr1.y);
                    }
                }
            }
            live -= live > 0;
        }
        // slide the window by 3 planes
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int i = 0; i < 8; ++i) win[q][i] = win[q][i + 3];
            win[q][8] = win[q][9] = win[q][10] = make_float2(0.f, 0.f);
        }
    }
    if (inside) {
        float* dst = a.proj + ((size_t)a.ch[c0] * a.Y + y) * a.X + x;
        const float b0[8] = {best0[0].x, best0[0].y, best0[1].x, best0[1].y, best0[2].x, best0[2].y, best0[3].x, best0[3].y};
        if (x + 7 < a.X && (a.X & 3) == 0) {
            reinterpret_cast<float4*>(dst)[0] = make_float4(b0[0], b0[1], b0[2], b0[3]);
            reinterpret_cast<float4*>(dst)[1] = make_float4(b0[4], b0[5], b0[6], b0[7]);
        } else {
#pragma unroll
            for (int p = 0; p < kBandPix; ++p)
                if (x + p < a.X) dst[p] = b0[p];
        }
        if (TWO) {
            float* dst1 = a.proj + ((size_t)a.ch[c0 + 1] * a.Y + y) * a.X + x;
            const float b1[8] = {best1[0].x, best1[0].y, best1[1].x, best1[1].y, best1[2].x, best1[2].y, best1[3].x, best1[3].y};
#pragma unroll
            for (int p = 0; p < kBandPix; ++p)
                if (x + p < a.X) dst1[p] = b1[p];
        }
    }
}

// ---- K5-K7 fused, TMA generation --------------------------------------------------------------------
// band_project2_kernel keeps one plane (16 bytes per thread) of raw voxels in flight, so every plane of the walk
// costs a full HBM round trip, and it spends most of its issue slots on bookkeeping.  This kernel does the same
// arithmetic with:
//   * the voxels by TMA: as soon as the tile's plane range [zlo-4, zhi+4] is known the CTA asks for every plane of
//     it at once - one 64 x 32 x 1 (x 1 channel) box per plane (cp.async.bulk.tensor.4d over the (X, Y, Z, C) map
//     of the cropped stack) into a ring of 16 stages, completing on one mbarrier per half ring; the loads overlap
//     the mask building and the weighted max reads shared memory.  Deeper ranges refill half a ring at a time;
//   * the height-map tile as uint16 in shared memory, loaded with 16-byte vectors (interior tiles), range by
//     warp reductions; the indicator masks of FOUR consecutive planes come out of one sweep over it (ballots),
//     together with a per-warp "plane present" flag - no atomics, no present-plane bitmap;
//   * the pending-mask window as a 9-slot CIRCULAR register file: the plane loop switches on (t - zlo) mod 9 so
//     that every slot index is a compile-time constant - no window shifting (12 % of the old kernel's instructions);
//   * the x-blur lookup tables precomputed on the host; a warp skips the y pass of a plane none of its 20 window
//     rows sees (one word per 4 rows of the x-blurred plane says which 8-pixel groups are non-zero).
constexpr int kB3PlaneBytes = kBandTY * kBandTX * 2;          // 4096: one channel of one plane of the tile
constexpr int kB3Chunk = 4;                                    // planes whose masks are built per sweep

template <bool AIRY, bool TWO>
__global__ void __launch_bounds__(kB2Threads, 2) band_project3_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                      const BandArgs a) {
    constexpr int NST = TWO ? 8 : 16;                             // ring stages (planes in flight)
    constexpr int HALF = NST / 2;
    constexpr int STAGE = (TWO ? 2 : 1) * kB3PlaneBytes;
    extern __shared__ __align__(128) unsigned char b3_ring[];     // [NST][STAGE] + 2 mbarriers
    __shared__ __align__(16) uint16_t cz_s[kB2CH][kB2CW];
    __shared__ __align__(16) uint32_t rowmask[2][kB3Chunk][kB2CH][4];
    __shared__ __align__(16) float r_s[2][kB2CH][kBandTX];       // row layout: [half][group][4]
    __shared__ uint32_t rownz[2][12];                             // [4 rows][8 groups] bits: group holds a non-zero
    __shared__ __align__(16) uint32_t pres[2][8];                 // bit p: warp w saw plane p of the chunk
    __shared__ float lut_lo[512], lut_hi[256];
    __shared__ int zlo_s, zhi_s;

    chain_release();
    chain_wait();
    if (band_index_error(a)) return;            // the reference raises before projecting

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // use_list: a small grid walks the tiles band_project4_kernel left behind; otherwise one tile per CTA
    const int ntiles = a.use_list ? st_load(a.status, ST_WORK_COUNT) : 1;
    for (int lin = a.use_list ? (int)blockIdx.x : 0; lin < ntiles; lin += a.use_list ? (int)gridDim.x : 1) {
    int bx = blockIdx.x, by = blockIdx.y;
    if (a.use_list) {
        const int tile = __ldcg(a.worklist + lin);
        bx = tile % a.tiles_x;
        by = tile / a.tiles_x;
    }
    const int x0 = bx * kBandTX, y0 = by * kBandTY;
    const uint32_t ring_s = smem_u32(b3_ring);
    const uint32_t bar_s = ring_s + NST * STAGE;
    if (tid == 0) {
        zlo_s = INT32_MAX;
        zhi_s = INT32_MIN;
        mbar_init(bar_s, 1);
        mbar_init(bar_s + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    lut_lo[tid] = __ldg(a.lut + tid);
    lut_lo[256 + tid] = __ldg(a.lut + 256 + tid);
    lut_hi[tid] = __ldg(a.lut + 512 + tid);
    {
        int lo = INT32_MAX, hi = INT32_MIN;
        auto put = [&](int v) {
            if (a.shift != 0) v = min(max(v + a.shift, 0), a.Z);
            lo = min(lo, v);
            hi = max(hi, v);
            return (uint32_t)v;
        };
        const bool interior = a.zvec && x0 >= kBandHalo && x0 + kBandTX + kBandHalo <= a.X && y0 >= kBandHalo &&
                              y0 + kBandTY + kBandHalo <= a.Y;
        if (interior) {                      // 48 rows x 20 int4, no clamping
            const int4* base = reinterpret_cast<const int4*>(a.zmap + (size_t)(y0 - kBandHalo) * a.X + (x0 - kBandHalo));
            const size_t rstride = (size_t)a.X / 4;
            int4 vals[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = tid + k * kB2Threads;
                vals[k] = i < kB2CH * (kB2CW / 4) ? __ldg(base + (size_t)(i / (kB2CW / 4)) * rstride + i % (kB2CW / 4))
                                                  : make_int4(0, 0, 0, 0);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = tid + k * kB2Threads;
                if (i < kB2CH * (kB2CW / 4)) {
                    const uint32_t p0 = put(vals[k].x) | (put(vals[k].y) << 16);
                    const uint32_t p1 = put(vals[k].z) | (put(vals[k].w) << 16);
                    *reinterpret_cast<uint2*>(&cz_s[i / (kB2CW / 4)][(i % (kB2CW / 4)) * 4]) = make_uint2(p0, p1);
                }
            }
        } else {
            constexpr int kPer = (kB2CH * kB2CW + kB2Threads - 1) / kB2Threads;
            int vals[kPer];
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                const int i = tid + k * kB2Threads;
                const int yy = min(max(y0 - kBandHalo + i / kB2CW, 0), a.Y - 1);
                const int xx = min(max(x0 - kBandHalo + i % kB2CW, 0), a.X - 1);
                vals[k] = i < kB2CH * kB2CW ? __ldg(a.zmap + (size_t)yy * a.X + xx) : 0;
            }
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                const int i = tid + k * kB2Threads;
                if (i < kB2CH * kB2CW) cz_s[i / kB2CW][i % kB2CW] = (uint16_t)put(vals[k]);
            }
        }
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        __syncthreads();                         // zlo_s / zhi_s initialised
        if (lane == 0) {
            atomicMin(&zlo_s, lo);
            atomicMax(&zhi_s, hi);
        }
    }
    __syncthreads();
    const int zlo = zlo_s, zhi = zhi_s;
    // planes the walk multiplies: z = zlo-4 .. zhi+4 inside the stack; plane index i = z - zbeg
    const int zbeg = max(zlo - 4, 0), zend = min(zhi + 4, a.Z - 1);
    const int npl = zend - zbeg + 1;
    const int c0 = blockIdx.z * kBandMaxCh;
    const int ch0 = a.ch[c0], ch1 = a.ch[c0 + (TWO ? 1 : 0)];
    // warp 0: lane 0 arms the barrier of a half ring with its byte count, then one lane per plane asks the TMA unit
    auto issue_half = [&](int hh) {
        const int first = hh * HALF, n = min(HALF, npl - first);
        const uint32_t bar = bar_s + 8 * (hh & 1);
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)n * STAGE);
        __syncwarp();
        if (lane < n) {
            const int i = first + lane;
            const uint32_t dst = ring_s + (i % NST) * STAGE;
            tma_load_4d(dst, &tmap, x0, y0, zbeg + i, ch0, bar);
            if (TWO) tma_load_4d(dst + kB3PlaneBytes, &tmap, x0, y0, zbeg + i, ch1, bar);
        }
    };
    if (warp == 0) {
        issue_half(0);
        if (npl > HALF) issue_half(1);
    }

    const float one_val = lut_lo[511] + lut_hi[255];
    const int g = tid & 7, row = tid >> 3;                     // 8 pixels x0+8g .. +7 of row y0+row
    const int x = x0 + g * kBandPix, y = y0 + row;
    const bool inside = y < a.Y && x < a.X;
    const float nb = -((float)a.pedestal + 8388608.0f);
    const float2 nbias = make_float2(nb, nb);
    uint32_t magic = 0x4B000000u;
    asm volatile("" : "+r"(magic));
    const unsigned char* my_vox = b3_ring + row * (kBandTX * 2) + g * 16;

    // indicator masks of planes t0 .. t0+3 into rowmask[cbuf] (+ which of them this warp saw)
    auto build_masks = [&](int t0, int cbuf) {
        uint32_t anyp[kB3Chunk] = {0, 0, 0, 0};
#pragma unroll
        for (int rr = 0; rr < kB2CH / 8; ++rr) {
            const int r = warp + 8 * rr;
            const int v0 = cz_s[r][lane], v1 = cz_s[r][32 + lane];
            const int v2 = lane < kB2CW - 64 ? cz_s[r][64 + lane] : -1;
            uint4 mine = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int p = 0; p < kB3Chunk; ++p) {
                const uint32_t b0 = __ballot_sync(0xffffffffu, v0 == t0 + p);
                const uint32_t b1 = __ballot_sync(0xffffffffu, v1 == t0 + p);
                const uint32_t b2 = __ballot_sync(0xffffffffu, v2 == t0 + p);
                anyp[p] |= b0 | b1 | b2;
                if (lane == p) mine = make_uint4(b0, b1, b2, 0u);
            }
            if (lane < kB3Chunk) *reinterpret_cast<uint4*>(&rowmask[cbuf][lane][r][0]) = mine;
        }
        if (lane == 0)
            pres[cbuf][warp] = (anyp[0] ? 1u : 0u) | (anyp[1] ? 2u : 0u) | (anyp[2] ? 4u : 0u) | (anyp[3] ? 8u : 0u);
    };
    // x pass of plane (cbuf, p): r_s[buf] and rownz[buf]
    auto x_pass = [&](int cbuf, int p, int buf) {
#pragma unroll
        for (int round = 0; round < 2; ++round) {
            const int r = round * 32 + row;
            if (round == 1 && r >= kB2CH) break;                 // warp-uniform (warps 0..3 take the second round)
            const uint4 m = *reinterpret_cast<const uint4*>(&rowmask[cbuf][p][r][0]);
            const uint32_t wlo = g < 4 ? m.x : m.y, whi = g < 4 ? m.y : m.z;
            const uint32_t W = __funnelshift_r(wlo, whi, 8 * (g & 3)) & 0xFFFFFFu;
            float o[8];
            if (W == 0u) {
#pragma unroll
                for (int q = 0; q < 8; ++q) o[q] = 0.f;
            } else if (W == 0xFFFFFFu) {
#pragma unroll
                for (int q = 0; q < 8; ++q) o[q] = one_val;
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) o[q] = lut_lo[(W >> q) & 511u] + lut_hi[(W >> (q + 9)) & 255u];
            }
            float4* dst = reinterpret_cast<float4*>(&r_s[buf][r][0]);
            dst[g] = make_float4(o[0], o[1], o[2], o[3]);
            dst[8 + g] = make_float4(o[4], o[5], o[6], o[7]);
            const uint32_t B = __ballot_sync(0xffffffffu, W != 0u);          // bit 8*rr + g, rows 4*warp + rr
            if (lane == 0) rownz[buf][round * 8 + warp] = B;
        }
    };

    float2 win[4][9];             // pending masks, slot of plane z = (z - (zlo - 4)) mod 9; two pixels per entry
    float2 best0[4], best1[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int i = 0; i < 9; ++i) win[q][i] = make_float2(0.f, 0.f);
        best0[q] = best1[q] = make_float2(0.f, 0.f);
    }
    int live = 0;

    build_masks(zlo, 0);
    __syncthreads();

    int buf = 0, slot = 0;
    uint32_t havebits = 0;
    for (int t = zlo; t <= zhi + 8; ++t) {
        const int rel = t - zlo, cbuf = (rel >> 2) & 1, pin = rel & 3;
        bool scatter = false;
        float2 an[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) an[q] = make_float2(0.f, 0.f);
        if (t <= zhi) {                                            // block-uniform
            // the chunk's flags are read once, right behind the barrier that published them: pres[cbuf] is rewritten
            // by the next chunk's first plane, which absent planes do not separate from this one by a barrier
            if (pin == 0) {
                const uint4 pa = *reinterpret_cast<const uint4*>(&pres[cbuf][0]);
                const uint4 pb = *reinterpret_cast<const uint4*>(&pres[cbuf][4]);
                havebits = pa.x | pa.y | pa.z | pa.w | pb.x | pb.y | pb.z | pb.w;
            }
            const bool have = (havebits >> pin) & 1u;
            if (have) {
                x_pass(cbuf, pin, buf);
                if (pin == 0 && t + kB3Chunk <= zhi) build_masks(t + kB3Chunk, cbuf ^ 1);
                __syncthreads();
                // y pass: a_new = sum_dy w2[dy] * r_s[row + dy][8g ..], skipped by warps whose 20 rows see only zeros
                const uint32_t* nzw = &rownz[buf][warp];
                if ((nzw[0] | nzw[1] | nzw[2] | nzw[3] | nzw[4]) != 0u) {
#pragma unroll
                    for (int dy = 0; dy < 17; ++dy) {
                        const float4* src = reinterpret_cast<const float4*>(&r_s[buf][row + dy][0]);
                        const float4 lo4 = src[g], hi4 = src[8 + g];
                        const float2 w = c_w2p[dy];
                        an[0] = __ffma2_rn(make_float2(lo4.x, lo4.y), w, an[0]);
                        an[1] = __ffma2_rn(make_float2(lo4.z, lo4.w), w, an[1]);
                        an[2] = __ffma2_rn(make_float2(hi4.x, hi4.y), w, an[2]);
                        an[3] = __ffma2_rn(make_float2(hi4.z, hi4.w), w, an[3]);
                    }
                    scatter = true;
                    live = 9;
                }
                buf ^= 1;
            } else if (pin == 0 && t + kB3Chunk <= zhi) {
                build_masks(t + kB3Chunk, cbuf ^ 1);
                __syncthreads();
            }
        }
        // voxels of plane z = t - 4 (all threads wait at the first plane of each half ring)
        const int z = t - 4;
        const bool walk = z >= zbeg && z <= zend;                  // block-uniform
        uint4 cur0 = make_uint4(0, 0, 0, 0), cur1 = cur0;
        int i = 0;
        if (walk) {
            i = z - zbeg;
            const int stage = i % NST;
            if (i % HALF == 0) mbar_wait(bar_s + 8 * ((i / HALF) & 1), (uint32_t)(i / NST) & 1u);
            if (live > 0) {
                cur0 = *reinterpret_cast<const uint4*>(my_vox + stage * STAGE);
                if (TWO) cur1 = *reinterpret_cast<const uint4*>(my_vox + stage * STAGE + kB3PlaneBytes);
            }
        }
        const bool mult = walk && live > 0;
        const bool edge = t - 4 < 4 || t + 4 > a.Z - 5;
        // plane t feeds the masks of z = t-4+k (slot (S+k) mod 9) with the (z, t) entry of the edge-replicating z
        // matrix (interior rows: the plain sigma=1 taps); then slot S = plane t-4 is complete: weighted max, clear
#define TSP_B3_CASE(S)                                                                                              \
    case S: {                                                                                                       \
        if (scatter) {                                                                                              \
            _Pragma("unroll") for (int k = 0; k < 9; ++k) {                                                         \
                float2 w = c_w1p[8 - k];                                                                            \
                if (edge) {                                                                                         \
                    const int zz = t - 4 + k;                                                                       \
                    const float we = (zz >= 0 && zz < a.Z) ? __ldg(a.wz + zz * 9 + (8 - k)) : 0.f;                  \
                    w = make_float2(we, we);                                                                        \
                }                                                                                                   \
                _Pragma("unroll") for (int q = 0; q < 4; ++q)                                                       \
                    win[q][(S + k) % 9] = __ffma2_rn(w, an[q], win[q][(S + k) % 9]);                                \
            }                                                                                                       \
        }                                                                                                           \
        if (mult) {                                                                                                 \
            const uint32_t w0[4] = {cur0.x, cur0.y, cur0.z, cur0.w};                                                \
            const uint32_t w1[4] = {cur1.x, cur1.y, cur1.z, cur1.w};                                                \
            _Pragma("unroll") for (int q = 0; q < 4; ++q) {                                                         \
                const float2 m = win[q][S];                                                                         \
                float2 f = __fadd2_rn(make_float2(band_u16f_lo(w0[q], magic), band_u16f_hi(w0[q], magic)), nbias);  \
                if (AIRY) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); }                                         \
                const float2 pr = __fmul2_rn(f, m);                                                                 \
                best0[q].x = fmaxf(best0[q].x, pr.x);                                                               \
                best0[q].y = fmaxf(best0[q].y, pr.y);                                                               \
                if (TWO) {                                                                                          \
                    float2 f1 = __fadd2_rn(make_float2(band_u16f_lo(w1[q], magic), band_u16f_hi(w1[q], magic)), nbias); \
                    if (AIRY) { f1.x = fmaxf(f1.x, 0.f); f1.y = fmaxf(f1.y, 0.f); }                                 \
                    const float2 pr1 = __fmul2_rn(f1, m);                                                           \
                    best1[q].x = fmaxf(best1[q].x, pr1.x);                                                          \
                    best1[q].y = fmaxf(best1[q].y, pr1.y);                                                          \
                }                                                                                                   \
            }                                                                                                       \
        }                                                                                                           \
        _Pragma("unroll") for (int q = 0; q < 4; ++q) win[q][S] = make_float2(0.f, 0.f);                            \
    } break;
        if (scatter || live > 0) {
            switch (slot) {
                TSP_B3_CASE(0) TSP_B3_CASE(1) TSP_B3_CASE(2) TSP_B3_CASE(3) TSP_B3_CASE(4)
                TSP_B3_CASE(5) TSP_B3_CASE(6) TSP_B3_CASE(7) TSP_B3_CASE(8)
            }
        }
#undef TSP_B3_CASE
        slot = slot == 8 ? 0 : slot + 1;
        live -= live > 0;
        // half a ring consumed and planes still to come: refill those stages
        if (walk && (i % HALF) == HALF - 1 && (i / HALF + 2) * HALF < npl) {
            __syncthreads();
            if (warp == 0) issue_half(i / HALF + 2);
        }
    }
    if (inside) {
        float* dst = a.proj + ((size_t)ch0 * a.Y + y) * a.X + x;
        reinterpret_cast<float4*>(dst)[0] = make_float4(best0[0].x, best0[0].y, best0[1].x, best0[1].y);
        reinterpret_cast<float4*>(dst)[1] = make_float4(best0[2].x, best0[2].y, best0[3].x, best0[3].y);
        if (TWO) {
            float* dst1 = a.proj + ((size_t)ch1 * a.Y + y) * a.X + x;
            reinterpret_cast<float4*>(dst1)[0] = make_float4(best1[0].x, best1[0].y, best1[1].x, best1[1].y);
            reinterpret_cast<float4*>(dst1)[1] = make_float4(best1[2].x, best1[2].y, best1[3].x, best1[3].y);
        }
    }
    __syncthreads();                             // the next tile re-initialises the barriers and the tile state
    if (tid == 0 && a.use_list) {
        asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar_s) : "memory");
        asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar_s + 8) : "memory");
    }
    }
}


// ---- K5-K7 fused, shallow-range kernel ---------------------------------------------------------------
// The sigma=30 score makes height maps smooth: in almost every 32 x 64 tile (with its 8-pixel halo) the heights span
// at most kB4NP planes.  For those tiles the pending-mask window of band_project3_kernel (72 registers, a switch per
// plane, a barrier-separated pipeline per plane) is not needed:
//   * the XY-blurred indicator planes A[zlo + j], j < kB4NP, are computed first and stay in registers (40);
//   * the walk over z = zlo-4 .. zhi+4 (at most kB4NW planes, all requested from the TMA unit at once, one mbarrier)
//     forms mask(z) = sum_j wz(z, zlo+j) A[j] on the fly - ascending j, the very FMA sequence of the scatter form, so
//     the results are bit-identical - multiplies and keeps the maximum; everything is unrolled against compile-time
//     (walk step, slot) pairs, the z weights come from a per-tile shared table (edge replication included);
//   * 80 registers and 73 KB of shared memory: three CTAs per SM instead of two.
// Tiles with a deeper range are appended to a worklist and done by band_project3_kernel (use_list) afterwards.
constexpr int kB4NP = 5;
constexpr int kB4NW = kB4NP + 8;
constexpr int kB4Smem = kB4NW * kB3PlaneBytes + 16;

template <bool AIRY>
__global__ void __launch_bounds__(kB2Threads, 3) band_project4_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                      const BandArgs a) {
    extern __shared__ __align__(128) unsigned char b4_ring[];     // [kB4NW][4096] + 1 mbarrier
    __shared__ __align__(16) float r_s[kB2CH][kBandTX];           // x-blurred plane; before that the uint16 height tile
    __shared__ __align__(16) uint32_t rowmask[kB4NP][kB2CH][4];
    __shared__ __align__(16) float2 wtab[kB4NW][kB4NP + 1];       // (w, w): weight of A[j] in mask(zlo - 4 + i)
    __shared__ float lut_lo[512], lut_hi[256];
    __shared__ uint32_t rownz[12];
    __shared__ uint32_t pres[8];
    __shared__ int zlo_s, zhi_s;
    uint16_t (*cz_s)[kB2CW] = reinterpret_cast<uint16_t (*)[kB2CW]>(&r_s[0][0]);

    chain_release();
    // the wait comes first: behind the table loads below it would expose their latency (the wait is a memory
    // barrier; the loads then cannot overlap the height-tile loads any more: +3 us measured)
    chain_wait();
    if (band_index_error(a)) return;            // the reference raises before projecting

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * kBandTX, y0 = blockIdx.y * kBandTY;
    const uint32_t ring_s = smem_u32(b4_ring);
    const uint32_t bar_s = ring_s + kB4NW * kB3PlaneBytes;
    if (tid == 0) {
        zlo_s = INT32_MAX;
        zhi_s = INT32_MIN;
        mbar_init(bar_s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    lut_lo[tid] = __ldg(a.lut + tid);
    lut_lo[256 + tid] = __ldg(a.lut + 256 + tid);
    lut_hi[tid] = __ldg(a.lut + 512 + tid);
    {
        int lo = INT32_MAX, hi = INT32_MIN;
        auto put = [&](int v) {
            if (a.shift != 0) v = min(max(v + a.shift, 0), a.Z);
            lo = min(lo, v);
            hi = max(hi, v);
            return (uint32_t)v;
        };
        const bool interior = a.zvec && x0 >= kBandHalo && x0 + kBandTX + kBandHalo <= a.X && y0 >= kBandHalo &&
                              y0 + kBandTY + kBandHalo <= a.Y;
        if (interior) {                      // 48 rows x 20 int4, no clamping
            const int4* base = reinterpret_cast<const int4*>(a.zmap + (size_t)(y0 - kBandHalo) * a.X + (x0 - kBandHalo));
            const size_t rstride = (size_t)a.X / 4;
            int4 vals[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = tid + k * kB2Threads;
                vals[k] = i < kB2CH * (kB2CW / 4) ? __ldg(base + (size_t)(i / (kB2CW / 4)) * rstride + i % (kB2CW / 4))
                                                  : make_int4(0, 0, 0, 0);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = tid + k * kB2Threads;
                if (i < kB2CH * (kB2CW / 4)) {
                    const uint32_t p0 = put(vals[k].x) | (put(vals[k].y) << 16);
                    const uint32_t p1 = put(vals[k].z) | (put(vals[k].w) << 16);
                    *reinterpret_cast<uint2*>(&cz_s[i / (kB2CW / 4)][(i % (kB2CW / 4)) * 4]) = make_uint2(p0, p1);
                }
            }
        } else {
            constexpr int kPer = (kB2CH * kB2CW + kB2Threads - 1) / kB2Threads;
            int vals[kPer];
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                const int i = tid + k * kB2Threads;
                const int yy = min(max(y0 - kBandHalo + i / kB2CW, 0), a.Y - 1);
                const int xx = min(max(x0 - kBandHalo + i % kB2CW, 0), a.X - 1);
                vals[k] = i < kB2CH * kB2CW ? __ldg(a.zmap + (size_t)yy * a.X + xx) : 0;
            }
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                const int i = tid + k * kB2Threads;
                if (i < kB2CH * kB2CW) cz_s[i / kB2CW][i % kB2CW] = (uint16_t)put(vals[k]);
            }
        }
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        __syncthreads();                         // zlo_s / zhi_s initialised
        if (lane == 0) {
            atomicMin(&zlo_s, lo);
            atomicMax(&zhi_s, hi);
        }
    }
    __syncthreads();
    const int zlo = zlo_s, zhi = zhi_s;
    if (zhi - zlo >= kB4NP) {                    // block-uniform: a deep tile goes to the generic kernel
        if (tid == 0 && blockIdx.z == 0) {
            const int slot = atomicAdd(const_cast<int32_t*>(a.status) + ST_WORK_COUNT, 1);
            a.worklist[slot] = blockIdx.y * gridDim.x + blockIdx.x;
        }
        return;
    }
    // planes the walk multiplies: z = zlo-4 .. zhi+4 inside the stack, all in flight at once
    const int zbeg = max(zlo - 4, 0), zend = min(zhi + 4, a.Z - 1);
    const int npl = zend - zbeg + 1;
    const int ch = a.ch[blockIdx.z];
    if (warp == 0) {
        if (lane == 0) mbar_expect_tx(bar_s, (uint32_t)npl * kB3PlaneBytes);
        __syncwarp();
        if (lane < npl) tma_load_4d(ring_s + lane * kB3PlaneBytes, &tmap, x0, y0, zbeg + lane, ch, bar_s);
    }
    if (tid < kB4NW * kB4NP) {                   // z weights: row z of the edge-replicating matrix, column zlo + j
        const int i = tid / kB4NP, j = tid % kB4NP;
        const int z = zlo - 4 + i, k = j - i + 8;
        const float w = (z >= 0 && z < a.Z && k >= 0 && k <= 8) ? __ldg(a.wz + z * 9 + k) : 0.f;
        wtab[i][j] = make_float2(w, w);
    }

    const float one_val = lut_lo[511] + lut_hi[255];
    const int g = tid & 7, row = tid >> 3;                     // 8 pixels x0+8g .. +7 of row y0+row
    const int x = x0 + g * kBandPix, y = y0 + row;
    const bool inside = y < a.Y && x < a.X;

    // indicator masks of planes zlo .. zlo+4 in one sweep over the height tile (+ which of them this warp saw)
    {
        uint32_t seen = 0;
#pragma unroll
        for (int rr = 0; rr < kB2CH / 8; ++rr) {
            const int r = warp + 8 * rr;
            const int v0 = cz_s[r][lane] - zlo, v1 = cz_s[r][32 + lane] - zlo;
            const int v2 = lane < kB2CW - 64 ? cz_s[r][64 + lane] - zlo : -1;
            uint4 mine = make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int p = 0; p < kB4NP; ++p) {
                const uint32_t b0 = __ballot_sync(0xffffffffu, v0 == p);
                const uint32_t b1 = __ballot_sync(0xffffffffu, v1 == p);
                const uint32_t b2 = __ballot_sync(0xffffffffu, v2 == p);
                if ((b0 | b1 | b2) != 0u) seen |= 1u << p;
                if (lane == p) mine = make_uint4(b0, b1, b2, 0u);
            }
            if (lane < kB4NP) *reinterpret_cast<uint4*>(&rowmask[lane][r][0]) = mine;
        }
        if (lane == 0) pres[warp] = seen;
    }
    __syncthreads();                             // masks and flags published; the height tile is dead from here on
    uint32_t havebits;
    {
        const uint4 pa = *reinterpret_cast<const uint4*>(&pres[0]);
        const uint4 pb = *reinterpret_cast<const uint4*>(&pres[4]);
        havebits = pa.x | pa.y | pa.z | pa.w | pb.x | pb.y | pb.z | pb.w;
    }

    float2 A[kB4NP][4];
    bool first = true;
#pragma unroll
    for (int j = 0; j < kB4NP; ++j) {
#pragma unroll
        for (int q = 0; q < 4; ++q) A[j][q] = make_float2(0.f, 0.f);
        if (!((havebits >> j) & 1u)) continue;                    // block-uniform
        if (!first) __syncthreads();                              // the previous plane's y pass is done with r_s
        first = false;
        // x pass: 17-tap blur of a binary row = two table lookups per pixel
#pragma unroll
        for (int round = 0; round < 2; ++round) {
            const int r = round * 32 + row;
            if (round == 1 && r >= kB2CH) break;                  // warp-uniform (warps 0..3 take the second round)
            const uint4 m = *reinterpret_cast<const uint4*>(&rowmask[j][r][0]);
            const uint32_t wlo = g < 4 ? m.x : m.y, whi = g < 4 ? m.y : m.z;
            const uint32_t W = __funnelshift_r(wlo, whi, 8 * (g & 3)) & 0xFFFFFFu;
            float o[8];
            if (W == 0u) {
#pragma unroll
                for (int q = 0; q < 8; ++q) o[q] = 0.f;
            } else if (W == 0xFFFFFFu) {
#pragma unroll
                for (int q = 0; q < 8; ++q) o[q] = one_val;
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) o[q] = lut_lo[(W >> q) & 511u] + lut_hi[(W >> (q + 9)) & 255u];
            }
            float4* dst = reinterpret_cast<float4*>(&r_s[r][0]);
            dst[g] = make_float4(o[0], o[1], o[2], o[3]);
            dst[8 + g] = make_float4(o[4], o[5], o[6], o[7]);
            const uint32_t B = __ballot_sync(0xffffffffu, W != 0u);
            if (lane == 0) rownz[round * 8 + warp] = B;
        }
        __syncthreads();
        // y pass, skipped by warps whose 20 rows of the x-blurred plane are all zero
        const uint32_t* nzw = &rownz[warp];
        if ((nzw[0] | nzw[1] | nzw[2] | nzw[3] | nzw[4]) != 0u) {
#pragma unroll
            for (int dy = 0; dy < 17; ++dy) {
                const float4* src = reinterpret_cast<const float4*>(&r_s[row + dy][0]);
                const float4 lo4 = src[g], hi4 = src[8 + g];
                const float2 w = c_w2p[dy];
                A[j][0] = __ffma2_rn(make_float2(lo4.x, lo4.y), w, A[j][0]);
                A[j][1] = __ffma2_rn(make_float2(lo4.z, lo4.w), w, A[j][1]);
                A[j][2] = __ffma2_rn(make_float2(hi4.x, hi4.y), w, A[j][2]);
                A[j][3] = __ffma2_rn(make_float2(hi4.z, hi4.w), w, A[j][3]);
            }
        }
    }

    const float nb = -((float)a.pedestal + 8388608.0f);
    const float2 nbias = make_float2(nb, nb);
    uint32_t magic = 0x4B000000u;
    asm volatile("" : "+r"(magic));
    const unsigned char* my_vox = b4_ring + row * (kBandTX * 2) + g * 16 - (size_t)zbeg * kB3PlaneBytes;
    float2 best[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) best[q] = make_float2(0.f, 0.f);
    mbar_wait(bar_s, 0);
    // The walk, specialised on the number of planes the tile spans (block-uniform switch): planes beyond zhi hold
    // exact zeros, their terms are left out at compile time (a tile spanning two planes runs 18 instead of 57
    // (step, plane) terms).  Dropping fma(w, 0, m) terms does not change m: bit-identical to the full form.
    auto walk = [&](auto span_tag) {
        constexpr int NR = decltype(span_tag)::value;
#pragma unroll
        for (int i = 0; i < NR + 8; ++i) {
            const int z = zlo - 4 + i;
            if (z > zend) break;                                   // block-uniform
            if (z < zbeg) continue;
            const uint4 cur = *reinterpret_cast<const uint4*>(my_vox + (size_t)z * kB3PlaneBytes);
            float2 m[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) m[q] = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                if (j > i || j < i - 8) continue;                  // outside the 9-band: compile time
                const float2 w = wtab[i][j];
#pragma unroll
                for (int q = 0; q < 4; ++q) m[q] = __ffma2_rn(w, A[j][q], m[q]);
            }
            const uint32_t w0[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float2 f = __fadd2_rn(make_float2(band_u16f_lo(w0[q], magic), band_u16f_hi(w0[q], magic)), nbias);
                if (AIRY) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); }
                const float2 pr = __fmul2_rn(f, m[q]);
                best[q].x = fmaxf(best[q].x, pr.x);
                best[q].y = fmaxf(best[q].y, pr.y);
            }
        }
    };
    switch (zhi - zlo + 1) {
        case 1: walk(std::integral_constant<int, 1>{}); break;
        case 2: walk(std::integral_constant<int, 2>{}); break;
        case 3: walk(std::integral_constant<int, 3>{}); break;
        case 4: walk(std::integral_constant<int, 4>{}); break;
        default: walk(std::integral_constant<int, 5>{}); break;
    }
    if (inside) {
        float* dst = a.proj + ((size_t)ch * a.Y + y) * a.X + x;
        reinterpret_cast<float4*>(dst)[0] = make_float4(best[0].x, best[0].y, best[1].x, best[1].y);
        reinterpret_cast<float4*>(dst)[1] = make_float4(best[2].x, best[2].y, best[3].x, best[3].y);
    }
}

constexpr int kB3Smem1 = 16 * kB3PlaneBytes + 16;
constexpr int kB3Smem2 = 8 * 2 * kB3PlaneBytes + 16;

// (X, Y, Z, C) map of the cropped stack, boxes of 64 x 32 x 1 x 1, zero fill outside
static int make_band_tensor_map(const uint16_t* base, size_t channel_stride, int C, int Z, int Y, int X,
                                CUtensorMap* out) {
    EncodeTiledFn encode = nullptr;
    int rc = get_tensor_map_encoder(&encode);
    if (rc) return rc;
    const cuuint64_t dims[4] = {(cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)Z, (cuuint64_t)C};
    const cuuint64_t strides[3] = {(cuuint64_t)X * 2, (cuuint64_t)X * Y * 2, (cuuint64_t)channel_stride * 2};
    const cuuint32_t box[4] = {(cuuint32_t)kBandTX, (cuuint32_t)kBandTY, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, (void*)base, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for the band stage C=%d Z=%d Y=%d X=%d", (int)r, C, Z, Y, X);
        return TSP_ERR_CUDA;
    }
    return TSP_OK;
}

// x-blur lookup tables of band_project3_kernel: the sigma=2 response of a binary row, split into taps 0..8 and 9..16
// (float sums in tap order, as the kernels that build them on the fly do)
static int get_band_lut(tsp_handle* h, const float** out) {
    std::lock_guard<std::mutex> lock(h->mu);
    auto it = h->tables.find("band_lut");
    if (it != h->tables.end()) {
        *out = (const float*)it->second;
        return TSP_OK;
    }
    std::vector<double> w2 = gaussian_taps(2.0);
    std::vector<float> tab(768);
    for (int i = 0; i < 512; ++i) {
        float s = 0.f;
        for (int k = 0; k < 9; ++k) s += ((i >> k) & 1) ? (float)w2[k] : 0.f;
        tab[i] = s;
    }
    for (int i = 0; i < 256; ++i) {
        float s = 0.f;
        for (int k = 0; k < 8; ++k) s += ((i >> k) & 1) ? (float)w2[9 + k] : 0.f;
        tab[512 + i] = s;
    }
    float* d = nullptr;
    TSP_CUDA(cudaMalloc(&d, tab.size() * sizeof(float)));
    TSP_CUDA(cudaMemcpy(d, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
    h->tables["band_lut"] = d;
    *out = d;
    return TSP_OK;
}

static int get_wz_table(tsp_handle* h, int Z, const float** out) {
    char key[32];
    snprintf(key, sizeof key, "wz%d", Z);
    std::lock_guard<std::mutex> lock(h->mu);
    if (!h->band_consts) {
        std::vector<double> w2 = gaussian_taps(2.0);
        float w2f[17];
        for (int i = 0; i < 17; ++i) w2f[i] = (float)w2[i];
        TSP_CUDA(cudaMemcpyToSymbol(c_w2, w2f, sizeof w2f));
        std::vector<double> w1 = gaussian_taps(1.0);
        float w1f[9];
        for (int i = 0; i < 9; ++i) w1f[i] = (float)w1[i];
        TSP_CUDA(cudaMemcpyToSymbol(c_w1, w1f, sizeof w1f));
        float2 w2p[17], w1p[9];
        for (int i = 0; i < 17; ++i) w2p[i] = make_float2(w2f[i], w2f[i]);
        for (int i = 0; i < 9; ++i) w1p[i] = make_float2(w1f[i], w1f[i]);
        TSP_CUDA(cudaMemcpyToSymbol(c_w2p, w2p, sizeof w2p));
        TSP_CUDA(cudaMemcpyToSymbol(c_w1p, w1p, sizeof w1p));
        h->band_consts = true;
    }
    auto it = h->tables.find(key);
    if (it != h->tables.end()) {
        *out = (const float*)it->second;
        return TSP_OK;
    }
    std::vector<double> wz = gaussian_taps(1.0);       // radius 4
    std::vector<float> tab((size_t)Z * 9, 0.f);
    for (int z = 0; z < Z; ++z) {
        std::vector<double> row(9, 0.0);
        for (int k = -4; k <= 4; ++k) {
            int zz = z + k;
            zz = zz < 0 ? 0 : (zz > Z - 1 ? Z - 1 : zz);
            row[zz - z + 4] += wz[k + 4];
        }
        for (int i = 0; i < 9; ++i) tab[(size_t)z * 9 + i] = (float)row[i];
    }
    float* d = nullptr;
    TSP_CUDA(cudaMalloc(&d, tab.size() * sizeof(float)));
    TSP_CUDA(cudaMemcpy(d, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
    h->tables[key] = d;
    *out = d;
    return TSP_OK;
}

__global__ void worklist_reset_kernel(int32_t* status) { status[ST_WORK_COUNT] = 0; }

static int launch_band_range(tsp_handle* h, const int32_t* d_zmap, int Z, int Y, int X, int shift,
                             int32_t* d_status, bool range_known, bool fused_check, cudaStream_t s,
                             const int32_t* d_zmap_other = nullptr) {
    if (range_known) {
        if (fused_check) return TSP_OK;          // the projection kernels apply the rule themselves
        band_check_kernel<<<1, 1, 0, s>>>(d_status, Z, shift, 1);
        TSP_LAUNCH_CHECK(h);
        return TSP_OK;
    }
    zmap_range_init_kernel<<<1, 1, 0, s>>>(d_status);
    TSP_LAUNCH_CHECK(h);
    const size_t n = (size_t)Y * X;
    size_t blocks = (n + 255) / 256;
    if (blocks > (size_t)h->sm_count * 8) blocks = (size_t)h->sm_count * 8;
    zmap_range_kernel<<<(int)blocks, 256, 0, s>>>(d_zmap, n, d_status);
    TSP_LAUNCH_CHECK(h);
    if (d_zmap_other) {                          // a separate map for the other channels: both must stay inside
        zmap_range_kernel<<<(int)blocks, 256, 0, s>>>(d_zmap_other, n, d_status);
        TSP_LAUNCH_CHECK(h);
    }
    band_check_kernel<<<1, 1, 0, s>>>(d_status, Z, shift, 0);
    TSP_LAUNCH_CHECK(h);
    return TSP_OK;
}

// channels are projected two per CTA; an odd channel count ends with a one-channel launch
static int launch_band_variant(tsp_handle* h, BandArgs a, dim3 grid, int pedestal, int C, cudaStream_t s) {
    // TMA path: rows 16-byte aligned (the tensor map's stride rule), tile not larger than the image
    const bool tma = a.vec && (a.X % 8 == 0) && a.X >= kBandTX && a.Y >= kBandTY && h->dbg.band_variant != 2;
    CUtensorMap tmap;
    if (tma) {
        int rc = make_band_tensor_map(a.stack + a.z0_offset, a.channel_stride, C, a.Z, a.Y, a.X, &tmap);
        if (rc) return rc;
        std::lock_guard<std::mutex> lock(h->mu);
        if (!h->band3_attr) {
            TSP_CUDA(cudaFuncSetAttribute(band_project3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kB3Smem1));
            TSP_CUDA(cudaFuncSetAttribute(band_project3_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kB3Smem1));
            TSP_CUDA(cudaFuncSetAttribute(band_project3_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kB3Smem2));
            TSP_CUDA(cudaFuncSetAttribute(band_project3_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kB3Smem2));
            h->band3_attr = true;
        }
    }
    // shallow tiles first, one channel per CTA; the deep ones come back through the worklist
    if (tma && a.worklist && h->dbg.band_variant != 3) {
        {
            std::lock_guard<std::mutex> lock(h->mu);
            if (!h->band4_attr) {
                TSP_CUDA(cudaFuncSetAttribute(band_project4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kB4Smem));
                TSP_CUDA(cudaFuncSetAttribute(band_project4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kB4Smem));
                h->band4_attr = true;
            }
        }
        dim3 g4(grid.x, grid.y, a.nch);
        if (pedestal) TSP_CUDA(launch_chained(band_project4_kernel<true>, g4, kB2Threads, kB4Smem, s, tmap, a));
        else TSP_CUDA(launch_chained(band_project4_kernel<false>, g4, kB2Threads, kB4Smem, s, tmap, a));
        TSP_LAUNCH_CHECK(h);
        a.use_list = 1;
    }
    // worklist mode: a small grid walks the (usually empty) list
    if (a.use_list) {
        const unsigned tiles = grid.x * grid.y, cap = 2u * (unsigned)h->sm_count;
        grid = dim3(tiles < cap ? tiles : cap, 1, 1);
    }
    const int pairs = a.nch / 2;
    if (pairs > 0) {
        BandArgs b = a;
        b.nch = 2 * pairs;
        dim3 g2(grid.x, grid.y, pairs);
        if (tma) {
            if (pedestal) TSP_CUDA(launch_chained(band_project3_kernel<true, true>, g2, kB2Threads, kB3Smem2, s, tmap, b));
            else TSP_CUDA(launch_chained(band_project3_kernel<false, true>, g2, kB2Threads, kB3Smem2, s, tmap, b));
        } else {
            if (pedestal) band_project2_kernel<true, true><<<g2, kB2Threads, 0, s>>>(b);
            else band_project2_kernel<false, true><<<g2, kB2Threads, 0, s>>>(b);
        }
    }
    if (a.nch & 1) {
        BandArgs b = a;
        b.ch[0] = a.ch[a.nch - 1];
        b.nch = 1;
        dim3 g1(grid.x, grid.y, 1);
        if (tma) {
            if (pedestal) TSP_CUDA(launch_chained(band_project3_kernel<true, false>, g1, kB2Threads, kB3Smem1, s, tmap, b));
            else TSP_CUDA(launch_chained(band_project3_kernel<false, false>, g1, kB2Threads, kB3Smem1, s, tmap, b));
        } else {
            if (pedestal) band_project2_kernel<true, false><<<g1, kB2Threads, 0, s>>>(b);
            else band_project2_kernel<false, false><<<g1, kB2Threads, 0, s>>>(b);
        }
    }
    return TSP_OK;
}

size_t band_worklist_bytes(int Y, int X) {
    const size_t tiles = (size_t)((X + kBandTX - 1) / kBandTX) * ((Y + kBandTY - 1) / kBandTY);
    return tiles * sizeof(int);
}

int launch_band_project_ex(tsp_handle* h, const uint16_t* d_stack, size_t channel_stride, size_t z0_offset,
                           const int32_t* d_zmap, float* d_proj, int C, int Z, int Y, int X, int ref_c,
                           int shift, int pedestal, int32_t* d_status, bool range_known, cudaStream_t s,
                           int* d_worklist, const int32_t* d_zmap_other) {
    // d_zmap_other: the height map of the non-reference channels when it is not clip(zmap + shift) (binned
    // manifold, SP:62-65); the shift is then already inside it
    if (d_zmap_other) {
        range_known = false;
        shift = 0;
    }
    if (Z > kBandMaxPlanes) {
        set_error("band projection supports at most %d planes", kBandMaxPlanes);
        return TSP_ERR_INVALID;
    }
    if (C > 17) {
        set_error("band projection supports at most 17 channels");
        return TSP_ERR_INVALID;
    }
    const float* wz = nullptr;
    int rc = get_wz_table(h, Z, &wz);
    if (rc) return rc;
    const float* lut = nullptr;
    rc = get_band_lut(h, &lut);
    if (rc) return rc;
    rc = launch_band_range(h, d_zmap, Z, Y, X, shift, d_status, range_known, range_known, s, d_zmap_other);
    if (rc) return rc;
    BandArgs a;
    a.worklist = d_worklist;
    a.use_list = 0;
    a.tiles_x = (X + kBandTX - 1) / kBandTX;
    a.fused_check = range_known ? 1 : 0;
    a.err_shift = shift;
    a.lut = lut;
    a.zvec = (X % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_zmap) & 15) == 0) ? 1 : 0;
    a.stack = d_stack;
    a.channel_stride = channel_stride;
    a.z0_offset = z0_offset;
    a.zmap = d_zmap;
    a.proj = d_proj;
    a.wz = wz;
    a.status = d_status;
    a.Z = Z;
    a.Y = Y;
    a.X = X;
    a.pedestal = pedestal;
    a.vec = (X % 8 == 0) && ((reinterpret_cast<uintptr_t>(d_stack) & 15) == 0) ? 1 : 0;
    dim3 grid((X + kBandTX - 1) / kBandTX, (Y + kBandTY - 1) / kBandTY, 1);
    // pass 1: every channel that uses the un-shifted mask (all of them when shift == 0)
    a.shift = 0;
    a.nch = 0;
    for (int c = 0; c < C; ++c)
        if ((shift == 0 && !d_zmap_other) || c == ref_c) a.ch[a.nch++] = c;
    grid.z = (a.nch + kBandMaxCh - 1) / kBandMaxCh;
    rc = launch_band_variant(h, a, grid, pedestal, C, s);
    if (rc) return rc;
    TSP_LAUNCH_CHECK(h);
    if ((shift != 0 || d_zmap_other) && C > 1) {
        if (d_zmap_other) a.zmap = d_zmap_other;
        if (d_worklist) {
            worklist_reset_kernel<<<1, 1, 0, s>>>(d_status);
            TSP_LAUNCH_CHECK(h);
        }
        a.shift = shift;
        a.nch = 0;
        for (int c = 0; c < C; ++c)
            if (c != ref_c) a.ch[a.nch++] = c;
        grid.z = (a.nch + kBandMaxCh - 1) / kBandMaxCh;
        rc = launch_band_variant(h, a, grid, pedestal, C, s);
        if (rc) return rc;
        TSP_LAUNCH_CHECK(h);
    }
    return TSP_OK;
}

// ---- bit-exact variant: materialise the one-hot volume and run the scipy-order passes ----------
__global__ void onehot_kernel(const int32_t* __restrict__ zmap, float* __restrict__ vol, int Z, size_t plane,
                              int shift, const int32_t* __restrict__ status) {
    if (st_load(status, ST_BAND_ERR)) return;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < plane; p += stride) {
        int v = zmap[p];
        if (shift != 0) v = min(max(v + shift, 0), Z);
        for (int z = 0; z < Z; ++z) vol[(size_t)z * plane + p] = z == v ? 1.f : 0.f;
    }
}

__global__ void mulmax_kernel(const uint16_t* __restrict__ chan, const float* __restrict__ mask,
                              float* __restrict__ proj, int Z, size_t plane, int pedestal,
                              const int32_t* __restrict__ status) {
    if (st_load(status, ST_BAND_ERR)) return;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < plane; p += stride) {
        float best = -INFINITY;
        for (int z = 0; z < Z; ++z) {
            int v = (int)chan[(size_t)z * plane + p] - pedestal;
            const float f = (float)(v > 0 ? v : 0);
            best = fmaxf(best, __fmul_rn(f, mask[(size_t)z * plane + p]));
        }
        proj[p] = best;
    }
}

int launch_band_project_bitexact_ex(tsp_handle* h, const uint16_t* d_stack, size_t channel_stride,
                                    size_t z0_offset, const int32_t* d_zmap, float* d_proj, int C, int Z,
                                    int Y, int X, int ref_c, int shift, int pedestal, float* d_volA,
                                    float* d_volB, int32_t* d_status, bool range_known, cudaStream_t s,
                                    const int32_t* d_zmap_other, const double* sigma_mask, bool fp64) {
    // Also the general form of the band stage: any sigma_mask (tsp_params), fp64 = scipy's summation order
    // (bit-exact), otherwise fp32 FMA accumulation.
    if (d_zmap_other) {
        range_known = false;
        shift = 0;
    }
    int rc = launch_band_range(h, d_zmap, Z, Y, X, shift, d_status, range_known, false, s, d_zmap_other);
    if (rc) return rc;
    const size_t plane = (size_t)Y * X;
    size_t blocks = (plane + 255) / 256;
    if (blocks > (size_t)h->sm_count * 32) blocks = (size_t)h->sm_count * 32;
    const double sig_ref[3] = {1.0, 2.0, 2.0};       // SP:70-71
    const double* sig = sigma_mask ? sigma_mask : sig_ref;
    for (int pass = 0; pass < 2; ++pass) {
        const int sh = pass == 0 ? 0 : shift;
        const bool two_maps = shift != 0 || d_zmap_other;
        if (pass == 1 && (!two_maps || C == 1)) break;
        onehot_kernel<<<(int)blocks, 256, 0, s>>>(pass == 1 && d_zmap_other ? d_zmap_other : d_zmap, d_volA, Z, plane, sh,
                                                  d_status);
        TSP_LAUNCH_CHECK(h);
        rc = gaussian_blur<float>(h, d_volA, d_volB, d_volA, Z, Y, X, sig, fp64, s);
        if (rc) return rc;
        for (int c = 0; c < C; ++c) {
            const bool uses = !two_maps ? (pass == 0) : ((c == ref_c) == (pass == 0));
            if (!uses) continue;
            mulmax_kernel<<<(int)blocks, 256, 0, s>>>(d_stack + (size_t)c * channel_stride + z0_offset, d_volB,
                                                      d_proj + (size_t)c * plane, Z, plane, pedestal, d_status);
            TSP_LAUNCH_CHECK(h);
        }
    }
    return TSP_OK;
}

// ---- output widening (the reference returns float64 / int64) ---------------------------------
__global__ void widen_kernel(const float* __restrict__ p32, const int32_t* __restrict__ z32,
                             double* __restrict__ p64, long long* __restrict__ z64, size_t np, size_t nz) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += stride) p64[i] = (double)p32[i];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nz; i += stride) z64[i] = (long long)z32[i];
}

int launch_widen_outputs(tsp_handle* h, const float* d_proj, const int32_t* d_zmap, double* d_proj64,
                         int64_t* d_zmap64, size_t nproj, size_t nz, cudaStream_t s) {
    size_t n = nproj > nz ? nproj : nz;
    size_t blocks = (n + 255) / 256;
    if (blocks > (size_t)h->sm_count * 16) blocks = (size_t)h->sm_count * 16;
    if (blocks == 0) blocks = 1;
    widen_kernel<<<(int)blocks, 256, 0, s>>>(d_proj, d_zmap, d_proj64, (long long*)d_zmap64, nproj, nz);
    TSP_LAUNCH_CHECK(h);
    return TSP_OK;
}

// ---- output narrowing for the movie driver: numpy's astype("uint16") of the float64 projection (BIM:481) and of the
// height map (SP:229-231) is a C cast: truncation toward zero (the values are non-negative and below 65536)
__global__ void narrow_kernel(const float* __restrict__ p32, const int32_t* __restrict__ z32,
                              uint16_t* __restrict__ p16, uint16_t* __restrict__ z16, size_t np, size_t nz) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += stride)
        p16[i] = (uint16_t)(long long)p32[i];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nz; i += stride) z16[i] = (uint16_t)z32[i];
}

int launch_narrow_outputs(tsp_handle* h, const float* d_proj, const int32_t* d_zmap, uint16_t* d_proj16,
                          uint16_t* d_zmap16, size_t nproj, size_t nz, cudaStream_t s) {
    size_t n = nproj > nz ? nproj : nz;
    size_t blocks = (n + 255) / 256;
    if (blocks > (size_t)h->sm_count * 16) blocks = (size_t)h->sm_count * 16;
    if (blocks == 0) blocks = 1;
    narrow_kernel<<<(int)blocks, 256, 0, s>>>(d_proj, d_zmap, d_proj16, d_zmap16, nproj, nz);
    TSP_LAUNCH_CHECK(h);
    return TSP_OK;
}

}  // namespace tsp
