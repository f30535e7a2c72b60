// Binned focus scores (reference surface_projection.py:39-53), their resampling back to the pixel grid
// (SP:59-65) and the continuous-manifold height map (SP:87-165).
//
//   * block_reduce_kernel: skimage.measure.block_reduce((1, b, b)) with np.mean / np.var - every plane is zero
//     padded to a multiple of b and each b x b block reduced (population variance, the padding counts);
//   * resize_argmax_kernel: skimage.transform.resize(score, (Z, Y, X)) for an upsampling, order 1 - which is
//     scipy.ndimage.zoom(order=1, mode='mirror', grid_mode=True): output o reads input (o + 0.5) * in / out - 0.5,
//     mirrored at the ends, two taps per axis, float64 arithmetic in scipy's order (value * wy * wx summed over
//     (y0,x0), (y0,x1), (y1,x0), (y1,x1)), float32 result - fused with the running first-maximum argmax over z, so
//     the resized volume is never written;
//   * resize_round_kernel: the same resampling of a coarse height map followed by np.round (half to even);
//   * manifold: the region growing of build_continues_manifold / find_pixel_plane.  The reference walks square
//     rings around the global score maximum; a cell depends on the cell visited just before it (same ring
//     segment) and on neighbours that are final before the segment starts.  One CTA processes a segment in
//     two phases: every thread turns one cell into a transition table T[v] = plane chosen if the previous
//     cell's plane is v (a handful of score reads per cell), then one thread chases the tables through shared
//     memory.  The result is the reference's, quirks included (row -1 wraps to the last row, SP:133-134;
//     two-apart neighbours give their truncated mean, SP:165).
#include "common.cuh"

namespace tsp {

// ---- block_reduce -----------------------------------------------------------------------------------
// one thread per (z, by, bx) block; float64 statistics, float32 result (numpy reduces float32 blocks in float32
// with pairwise sums - the difference is rounding, covered by the tolerance rule).  combine: 0 store, 1 multiply
// the stored value by the result (SP:51: atoh_score * zo_score).
__global__ void block_reduce_kernel(const float* __restrict__ vol, float* __restrict__ out, int Z, int Y, int X,
                                    int bin, int by_n, int bx_n, int variance, int combine) {
    const size_t n = (size_t)Z * by_n * bx_n;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const double cnt = (double)bin * (double)bin;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int bx = (int)(i % bx_n), by = (int)((i / bx_n) % by_n), z = (int)(i / ((size_t)bx_n * by_n));
        const float* plane = vol + (size_t)z * Y * X;
        const int y1 = min(by * bin + bin, Y), x1 = min(bx * bin + bin, X);
        double sum = 0.0;
        for (int y = by * bin; y < y1; ++y)
            for (int x = bx * bin; x < x1; ++x) sum += (double)plane[(size_t)y * X + x];
        const double mean = sum / cnt;
        double res = mean;
        if (variance) {
            double acc = 0.0;
            for (int y = by * bin; y < y1; ++y)
                for (int x = bx * bin; x < x1; ++x) {
                    const double d = (double)plane[(size_t)y * X + x] - mean;
                    acc += d * d;
                }
            const double pad = cnt - (double)(y1 - by * bin) * (double)(x1 - bx * bin);     // zero padding
            acc += pad * mean * mean;
            res = acc / cnt;
        }
        const float r = (float)res;
        out[i] = combine ? __fmul_rn(out[i], r) : r;
    }
}

int launch_block_reduce(tsp_handle* h, const float* d_vol, float* d_out, int Z, int Y, int X, int bin,
                        bool variance, bool multiply, cudaStream_t s) {
    const int by_n = (Y + bin - 1) / bin, bx_n = (X + bin - 1) / bin;
    const size_t n = (size_t)Z * by_n * bx_n;
    size_t blocks = (n + 127) / 128;
    if (blocks > (size_t)h->sm_count * 32) blocks = (size_t)h->sm_count * 32;
    block_reduce_kernel<<<(int)blocks, 128, 0, s>>>(d_vol, d_out, Z, Y, X, bin, by_n, bx_n, variance ? 1 : 0,
                                                    multiply ? 1 : 0);
    TSP_LAUNCH_CHECK(h);
    return TSP_OK;
}

// ---- order-1 resampling of scipy.ndimage.zoom(mode='mirror', grid_mode=True) ---------------------------
struct AxisTap {
    int i0, i1;
    double w0, w1;
};

__device__ __forceinline__ int mirror_index(int i, int n) {
    if (n <= 1) return 0;
    const int sz2 = 2 * n - 2;
    if (i < 0) i = -i;
    i %= sz2;
    return i > n - 1 ? sz2 - i : i;
}

__device__ __forceinline__ AxisTap zoom_tap(int o, int n_in, int n_out) {
    AxisTap t;
    if (n_in == n_out) {                     // zoom 1: scipy does not interpolate along this axis
        t.i0 = t.i1 = o;
        t.w0 = 1.0;
        t.w1 = 0.0;
        return t;
    }
    const double zi = (double)n_in / (double)n_out;
    double cc = __dadd_rn(__dmul_rn((double)o + 0.5, zi), -0.5);
    if (n_in <= 1) {
        cc = 0.0;
    } else {
        const int sz2 = 2 * n_in - 2;
        if (cc < 0.0) {
            cc = (double)sz2 * (double)(int)(-cc / (double)sz2) + cc;
            cc = cc <= (double)(1 - n_in) ? cc + (double)sz2 : -cc;
        } else if (cc > (double)(n_in - 1)) {
            cc -= (double)sz2 * (double)(int)(cc / (double)sz2);
            if (cc > (double)(n_in - 1)) cc = (double)sz2 - cc;
        }
    }
    const double fl = floor(cc);
    const int s = (int)fl;
    const double frac = cc - fl;
    t.i0 = mirror_index(s, n_in);
    t.i1 = mirror_index(s + 1, n_in);
    t.w0 = 1.0 - frac;
    t.w1 = frac;
    return t;
}

__device__ __forceinline__ float zoom_value(const float* __restrict__ plane, int cx, const AxisTap& ty,
                                            const AxisTap& tx, bool iy, bool ix) {
    const float* r0 = plane + (size_t)ty.i0 * cx;
    const float* r1 = plane + (size_t)ty.i1 * cx;
    // a non-interpolated axis contributes no factor and a single tap (scipy skips it)
    if (!iy && !ix) return r0[tx.i0];
    double t;
    if (iy && ix) {
        t = __dmul_rn(__dmul_rn((double)r0[tx.i0], ty.w0), tx.w0);
        t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)r0[tx.i1], ty.w0), tx.w1));
        t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)r1[tx.i0], ty.w1), tx.w0));
        t = __dadd_rn(t, __dmul_rn(__dmul_rn((double)r1[tx.i1], ty.w1), tx.w1));
    } else if (iy) {
        t = __dmul_rn((double)r0[tx.i0], ty.w0);
        t = __dadd_rn(t, __dmul_rn((double)r1[tx.i0], ty.w1));
    } else {
        t = __dmul_rn((double)r0[tx.i0], tx.w0);
        t = __dadd_rn(t, __dmul_rn((double)r0[tx.i1], tx.w1));
    }
    return (float)t;
}

// thread = one pixel; coarse score (Z, cy, cx) -> zmap (Y, X) = z_offset + first maximum over z of the resized score
__global__ void resize_argmax_kernel(const float* __restrict__ score, int32_t* __restrict__ zmap, int Z, int Y, int X,
                                     int cy, int cx, int z_offset, int32_t* __restrict__ status) {
    const size_t plane = (size_t)Y * X, cplane = (size_t)cy * cx;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    int zmin = INT32_MAX, zmax = INT32_MIN;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < plane; p += stride) {
        const int y = (int)(p / X), x = (int)(p % X);
        const AxisTap ty = zoom_tap(y, cy, Y), tx = zoom_tap(x, cx, X);
        float best = zoom_value(score, cx, ty, tx, cy != Y, cx != X);
        int bz = 0;
        for (int z = 1; z < Z; ++z) {
            const float v = zoom_value(score + (size_t)z * cplane, cx, ty, tx, cy != Y, cx != X);
            if (v > best) {
                best = v;
                bz = z;
            }
        }
        zmap[p] = bz + z_offset;
        zmin = min(zmin, bz + z_offset);
        zmax = max(zmax, bz + z_offset);
    }
    zmin = __reduce_min_sync(0xffffffffu, zmin);
    zmax = __reduce_max_sync(0xffffffffu, zmax);
    if ((threadIdx.x & 31) == 0 && zmax >= zmin) {
        atomicMax(&status[ST_ZMAX], zmax);
        atomicMax(&status[ST_ZMIN_INV], INT32_MAX - zmin);
    }
}

int launch_resize_argmax(tsp_handle* h, const float* d_score, int32_t* d_zmap, int Z, int Y, int X, int cy, int cx,
                         int z_offset, int32_t* d_status, cudaStream_t s) {
    const size_t plane = (size_t)Y * X;
    size_t blocks = (plane + 127) / 128;
    if (blocks > (size_t)h->sm_count * 32) blocks = (size_t)h->sm_count * 32;
    resize_argmax_kernel<<<(int)blocks, 128, 0, s>>>(d_score, d_zmap, Z, Y, X, cy, cx, z_offset, d_status);
    TSP_LAUNCH_CHECK(h);
    return TSP_OK;
}

// SP:63-65: np.round(resize(chosen_z.astype('float32'), (Y, X))).astype('int'); `shift` != 0 first applies
// np.clip(chosen_z + shift, 0, clip_hi) (SP:62) to the coarse map.
__global__ void resize_round_kernel(const int32_t* __restrict__ coarse, int32_t* __restrict__ zmap, int Y, int X, int cy,
                                    int cx, int shift, int clip_hi) {
    const size_t plane = (size_t)Y * X;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    auto val = [&](int r, int c) {
        int v = coarse[(size_t)r * cx + c];
        if (shift != 0) v = min(max(v + shift, 0), clip_hi);
        return (double)(float)v;
    };
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < plane; p += stride) {
        const int y = (int)(p / X), x = (int)(p % X);
        const AxisTap ty = zoom_tap(y, cy, Y), tx = zoom_tap(x, cx, X);
        const bool iy = cy != Y, ix = cx != X;
        double t;
        if (iy && ix) {
            t = __dmul_rn(__dmul_rn(val(ty.i0, tx.i0), ty.w0), tx.w0);
            t = __dadd_rn(t, __dmul_rn(__dmul_rn(val(ty.i0, tx.i1), ty.w0), tx.w1));
            t = __dadd_rn(t, __dmul_rn(__dmul_rn(val(ty.i1, tx.i0), ty.w1), tx.w0));
            t = __dadd_rn(t, __dmul_rn(__dmul_rn(val(ty.i1, tx.i1), ty.w1), tx.w1));
        } else if (iy) {
            t = __dadd_rn(__dmul_rn(val(ty.i0, tx.i0), ty.w0), __dmul_rn(val(ty.i1, tx.i0), ty.w1));
        } else if (ix) {
            t = __dadd_rn(__dmul_rn(val(ty.i0, tx.i0), tx.w0), __dmul_rn(val(ty.i0, tx.i1), tx.w1));
        } else {
            t = val(ty.i0, tx.i0);
        }
        zmap[p] = (int)rintf((float)t);              // np.round: half to even, on the float32 the resize returns
    }
}

int launch_resize_round(tsp_handle* h, const int32_t* d_coarse, int32_t* d_zmap, int Y, int X, int cy, int cx,
                        int shift, int clip_hi, cudaStream_t s) {
    const size_t plane = (size_t)Y * X;
    size_t blocks = (plane + 127) / 128;
    if (blocks > (size_t)h->sm_count * 32) blocks = (size_t)h->sm_count * 32;
    resize_round_kernel<<<(int)blocks, 128, 0, s>>>(d_coarse, d_zmap, Y, X, cy, cx, shift, clip_hi);
    TSP_LAUNCH_CHECK(h);
    return TSP_OK;
}

// ---- continuous manifold (SP:87-165) -----------------------------------------------------------------
__device__ __forceinline__ uint32_t ordered_bits(float v) {
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// np.argmax(score) over the flattened volume: largest value, first position.  key = (value, ~index)
__global__ void argmax3d_kernel(const float* __restrict__ score, size_t n, unsigned long long* __restrict__ best) {
    unsigned long long mine = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long key = ((unsigned long long)ordered_bits(score[i]) << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)i);
        mine = key > mine ? key : mine;
    }
    for (int o = 16; o; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, mine, o);
        mine = other > mine ? other : mine;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(best, mine);
}

constexpr int kMfThreads = 1024;
constexpr int kMfMaxPlanes = 254;
constexpr int kMfTableBytes = 160 * 1024;

struct MfCtx {
    const float* score;
    int32_t* chosen;
    int P, R, C;
    size_t plane;
};

// value of the neighbour in direction dir (0 up, 1 down, 2 left, 3 right) as find_pixel_plane sees it: -1 when the
// reference does not look there or the cell has no plane yet.  Row -1 wraps to the last row (SP:133-134).
__device__ __forceinline__ int mf_neighbour(const MfCtx& m, int row, int col, int dir) {
    if (dir == 0) {
        const int rr = row > 0 ? row - 1 : m.R - 1;
        return m.chosen[(size_t)rr * m.C + col];
    }
    if (dir == 1) return row < m.R - 1 ? m.chosen[(size_t)(row + 1) * m.C + col] : -1;
    if (dir == 2) return col > 0 ? m.chosen[(size_t)row * m.C + col - 1] : -1;
    return col < m.C - 1 ? m.chosen[(size_t)row * m.C + col + 1] : -1;
}

// first maximum of planes lo .. hi-1 (two or three of them, never more); the loads do not depend on each other: one
// trip to L2, not three
__device__ __forceinline__ int mf_argmax_window(const MfCtx& m, int row, int col, int lo, int hi) {
    const float* s = m.score + (size_t)row * m.C + col;
    const float v0 = s[(size_t)lo * m.plane];
    const float v1 = lo + 1 < hi ? s[(size_t)(lo + 1) * m.plane] : v0;
    const float v2 = lo + 2 < hi ? s[(size_t)(lo + 2) * m.plane] : v0;
    float best = v0;
    int bz = lo;
    if (v1 > best) {
        best = v1;
        bz = lo + 1;
    }
    if (v2 > best) bz = lo + 2;
    return bz;
}

// SP:153-165 for a resolved neighbour pair (second < 0: none)
__device__ __forceinline__ int mf_decide(const MfCtx& m, int row, int col, int first, int second) {
    if (second < 0 || first == second) return mf_argmax_window(m, row, col, max(0, first - 1), min(m.P, first + 2));
    const int d = first - second;
    if (d == 1 || d == -1) {
        const int lo = min(first, second);
        return mf_argmax_window(m, row, col, lo, min(m.P, lo + 2));
    }
    return (first + second) >> 1;            // float mean stored into the int map: truncation
}

// One ring segment: `n` cells starting at (r0, c0), stepping (dr, dc).  prev_dir = direction from a cell to the
// cell visited before it.
//   phase A (all threads): every cell's rule as a transition table T_k: plane of the predecessor -> plane of the cell;
//   phase B: the chain c_k = T_k[c_(k-1)].  Function composition is associative, so the chain is cut into 32 runs:
//            warp w first sends EVERY possible incoming plane through its run (lane = plane, independent lookups:
//            the composite H_w of its tables), then one lane walks H_0 .. H_(w-1) from the segment's start to find
//            the plane that really enters run w and follows its own run with it.  Depth n/16 + 31 dependent
//            shared-memory lookups instead of n (a 2048-cell edge: 159 instead of 2048); short segments keep the
//            single walk;
//   phase C (all threads): publish.
constexpr int kMfSeqCells = 96;              // segments up to this length are walked by one thread
constexpr int kMfCorners = 32;               // cells per chunk whose table is filled by a whole warp

__device__ void mf_segment(const MfCtx& m, int r0, int c0, int dr, int dc, int n, int prev_dir, unsigned char* tab,
                           unsigned char* outv, unsigned char* comp, int* corners, int ppad, int chunk_cells) {
    const int tid = threadIdx.x;
    int& corner_n = corners[0];
    int* corner_k = corners + 1;
    for (int base = 0; base < n; base += chunk_cells) {
        const int cn = min(chunk_cells, n - base);
        // phase A: transition tables
        for (int k = tid; k < cn; k += kMfThreads) {
            const int row = r0 + (base + k) * dr, col = c0 + (base + k) * dc;
            unsigned char* T = tab + (size_t)k * ppad;
            int nb[4];
#pragma unroll
            for (int d = 0; d < 4; ++d) nb[d] = mf_neighbour(m, row, col, d);
            const int pd = k == 0 ? -1 : prev_dir;         // the first cell of a chunk finds its predecessor in memory
            // the first two neighbours with a plane, in the reference's order; the predecessor counts as set
            int e0 = -2, e1 = -2;                           // -2 none, -3 predecessor, >= 0 a plane from memory
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const int v = d == pd ? -3 : (nb[d] >= 0 ? nb[d] : -2);
                if (v == -2) continue;
                if (e0 == -2) e0 = v;
                else if (e1 == -2) e1 = v;
            }
            // tables are written four entries per 32-bit store; consecutive cells are an odd number of words apart
            // (ppad / 4 is odd), so the stores of a warp fall into 32 different banks (byte stores at a pitch of
            // 64 bytes were a 16-way bank conflict: 33 us of a 2048-cell segment)
            uint32_t* T4 = reinterpret_cast<uint32_t*>(T);
            if (e0 != -3 && e1 != -3) {                     // the predecessor is not consulted: a constant
                const int r = e0 == -2 ? 0 : mf_decide(m, row, col, e0, e1 == -2 ? -1 : e1);
                const uint32_t w = 0x01010101u * (uint32_t)r;
                for (int v = 0; v < ppad; v += 4) T4[v >> 2] = w;
            } else if (e1 == -2) {                          // only the predecessor: three-plane window around it
                // (the corner cells of a ring.)  One entry per plane, each an argmax over three planes of the score:
                // left to this thread they are P dependent trips to L2 (23 us at 64 planes, with the other 1023
                // threads waiting at the barrier) - the cell goes on a list and a whole warp fills its table below
                const int slot = atomicAdd(&corner_n, 1);
                if (slot < kMfCorners) {
                    corner_k[slot] = k;
                } else {
                    for (int v = 0; v < m.P; v += 4) {
                        uint32_t w = 0;
                        for (int i = 0; i < 4 && v + i < m.P; ++i) w |= (uint32_t)mf_decide(m, row, col, v + i, -1) << (8 * i);
                        T4[v >> 2] = w;
                    }
                }
            } else {                                        // the predecessor and one plane from memory
                // the three rules that can fire (SP:153-163) look at planes a-1, a, a+1 only: one round of loads
                const int a = e0 == -3 ? e1 : e0;
                const float* sp = m.score + (size_t)row * m.C + col;
                const float s0 = sp[(size_t)a * m.plane];
                const float sm = a > 0 ? sp[(size_t)(a - 1) * m.plane] : s0;
                const float s1 = a + 1 < m.P ? sp[(size_t)(a + 1) * m.plane] : s0;
                int same = a > 0 ? a - 1 : a;                              // mf_decide(a, a): planes a-1 .. a+1
                {
                    float best = a > 0 ? sm : s0;
                    if (a > 0 && s0 > best) { best = s0; same = a; }
                    if (a + 1 < m.P && s1 > best) same = a + 1;
                }
                const int below = a > 0 ? (s0 > sm ? a : a - 1) : 0;       // mf_decide(a - 1, a): planes a-1, a
                const int above = a + 1 < m.P ? (s1 > s0 ? a + 1 : a) : 0; // mf_decide(a + 1, a): planes a, a+1
                for (int v = 0; v < m.P; v += 4) {
                    uint32_t w = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int u = v + i;
                        const int t = u == a ? same : u == a - 1 ? below : u == a + 1 ? above : (u + a) >> 1;
                        w |= (uint32_t)(t & 0xff) << (8 * i);
                    }
                    T4[v >> 2] = w;
                }
            }
        }
        __syncthreads();
        {   // tables of the listed corner cells: warp per cell, lane per plane
            const int ncorner = min(corner_n, kMfCorners);
            if (ncorner > 0) {                              // block-uniform
                for (int e = tid >> 5; e < ncorner; e += kMfThreads / 32) {
                    const int k = corner_k[e];
                    const int row = r0 + (base + k) * dr, col = c0 + (base + k) * dc;
                    for (int v = tid & 31; v < m.P; v += 32) tab[(size_t)k * ppad + v] = (unsigned char)mf_decide(m, row, col, v, -1);
                }
                __syncthreads();
                if (tid == 0) corner_n = 0;
            }
        }
        // phase B (the first table of a chunk is constant: whatever enters it, 0 here, is ignored)
        if (cn <= kMfSeqCells) {
            if (tid == 0) {
                int c = 0;
                for (int k = 0; k < cn; ++k) {
                    c = tab[(size_t)k * ppad + c];
                    outv[k] = (unsigned char)c;
                }
            }
        } else {
            const int lane = tid & 31, warp = tid >> 5;
            const int per = (cn + 31) / 32;                  // cells per run
            const int k0 = min(warp * per, cn), k1 = min(k0 + per, cn);
            {   // H_warp[v] for every plane v: lane takes v = lane, lane + 32, ... (eight independent chains)
                int x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = min(lane + 32 * i, m.P - 1);
                for (int k = k0; k < k1; ++k) {
                    const unsigned char* T = tab + (size_t)k * ppad;
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = T[x[i]];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (lane + 32 * i < m.P) comp[warp * ppad + lane + 32 * i] = (unsigned char)x[i];
            }
            __syncthreads();
            if (lane == 0 && k0 < k1) {
                int c = 0;
                for (int u = 0; u < warp; ++u) c = comp[u * ppad + c];       // what enters this run
                for (int k = k0; k < k1; ++k) {
                    c = tab[(size_t)k * ppad + c];
                    outv[k] = (unsigned char)c;
                }
            }
        }
        __syncthreads();
        // phase C: publish
        for (int k = tid; k < cn; k += kMfThreads) {
            const int row = r0 + (base + k) * dr, col = c0 + (base + k) * dc;
            m.chosen[(size_t)row * m.C + col] = outv[k];
        }
        __threadfence_block();
        __syncthreads();
    }
}

// one cell evaluated straight from memory (used where a cell depends on an earlier cell of its own segment other
// than its predecessor: the wrapped row of a full-height left edge)
__device__ void mf_single(const MfCtx& m, int row, int col) {
    if (threadIdx.x == 0) {
        int e0 = -1, e1 = -1;
        for (int d = 0; d < 4; ++d) {
            const int v = mf_neighbour(m, row, col, d);
            if (v < 0) continue;
            if (e0 < 0) e0 = v;
            else if (e1 < 0) e1 = v;
        }
        m.chosen[(size_t)row * m.C + col] = e0 < 0 ? 0 : mf_decide(m, row, col, e0, e1);
    }
    __threadfence_block();
    __syncthreads();
}

__global__ void __launch_bounds__(kMfThreads, 1)
manifold_kernel(const float* __restrict__ score, int32_t* chosen, int P, int R, int C,
                const unsigned long long* __restrict__ best, int32_t* __restrict__ status, int ppad, int chunk_cells) {
    extern __shared__ __align__(16) unsigned char mf_smem[];
    unsigned char* tab = mf_smem;
    unsigned char* outv = mf_smem + (size_t)chunk_cells * ppad;
    unsigned char* comp = outv + chunk_cells;                    // [32 runs][ppad]: composite table of every run
    __shared__ int corners[1 + kMfCorners];                      // [0]: cells on the list, then their indices
    if (threadIdx.x == 0) corners[0] = 0;
    MfCtx m;
    m.score = score;
    m.chosen = chosen;
    m.P = P;
    m.R = R;
    m.C = C;
    m.plane = (size_t)R * C;
    for (size_t i = threadIdx.x; i < m.plane; i += kMfThreads) chosen[i] = -1;
    const uint32_t flat = 0xFFFFFFFFu - (uint32_t)(*best & 0xFFFFFFFFull);
    const int p0 = (int)(flat / m.plane), r0 = (int)((flat % m.plane) / C), c0 = (int)(flat % C);
    __syncthreads();
    if (threadIdx.x == 0) chosen[(size_t)r0 * C + c0] = p0;
    __threadfence_block();
    __syncthreads();
    const int reach = max(max(c0, r0), max(C - 1 - c0, R - 1 - r0));
    for (int d = 1; d <= reach; ++d) {
        // right edge, lower half: rows r0 .. r0+d going down
        if (c0 + d < C) {
            const int last = min(r0 + d, R - 1);
            mf_segment(m, r0, c0 + d, 1, 0, last - r0 + 1, 0, tab, outv, comp, corners, ppad, chunk_cells);
        }
        // bottom edge: columns c0+d-1 .. c0-d going left
        if (r0 + d < R) {
            const int first = min(c0 + d - 1, C - 1), last = max(c0 - d, 0);
            if (first >= last) mf_segment(m, r0 + d, first, 0, -1, first - last + 1, 3, tab, outv, comp, corners, ppad, chunk_cells);
        }
        // left edge: rows r0+d-1 .. r0-d going up
        if (c0 - d >= 0) {
            const int first = min(r0 + d - 1, R - 1), last = max(r0 - d, 0);
            if (first >= last) {
                // a segment that holds both the last row and row 0: row 0 looks at the last row (wrap), which is
                // an earlier cell of this very segment - finish the segment at row 1 and do row 0 on its own
                const bool wrap = last == 0 && first == R - 1 && R > 1;
                const int n = first - last + 1 - (wrap ? 1 : 0);
                if (n > 0) mf_segment(m, first, c0 - d, -1, 0, n, 1, tab, outv, comp, corners, ppad, chunk_cells);
                if (wrap) mf_single(m, 0, c0 - d);
            }
        }
        // top edge: columns c0-d+1 .. c0+d going right
        if (r0 - d >= 0) {
            const int first = max(c0 - d + 1, 0), last = min(c0 + d, C - 1);
            if (first <= last) mf_segment(m, r0 - d, first, 0, 1, last - first + 1, 2, tab, outv, comp, corners, ppad, chunk_cells);
        }
        // right edge, upper half: rows r0-d+1 .. r0-1 going down
        if (c0 + d < C) {
            const int first = max(r0 - d + 1, 0), last = r0 - 1;
            if (first <= last) mf_segment(m, first, c0 + d, 1, 0, last - first + 1, 0, tab, outv, comp, corners, ppad, chunk_cells);
        }
    }
    // height-map range for the band stage's IndexError rule
    int zmin = INT32_MAX, zmax = INT32_MIN;
    for (size_t i = threadIdx.x; i < m.plane; i += kMfThreads) {
        const int v = chosen[i];
        zmin = min(zmin, v);
        zmax = max(zmax, v);
    }
    zmin = __reduce_min_sync(0xffffffffu, zmin);
    zmax = __reduce_max_sync(0xffffffffu, zmax);
    if ((threadIdx.x & 31) == 0 && zmax >= zmin) {
        atomicMax(&status[ST_ZMAX], zmax);
        atomicMax(&status[ST_ZMIN_INV], INT32_MAX - zmin);
    }
}

size_t manifold_scratch_bytes() { return 256; }

// d_chosen (R, C) int32 = build_continues_manifold(score (P, R, C)); d_scratch: manifold_scratch_bytes()
int launch_manifold(tsp_handle* h, const float* d_score, int32_t* d_chosen, int P, int R, int C, int32_t* d_status,
                    void* d_scratch, cudaStream_t s) {
    if (P > kMfMaxPlanes) {
        set_error("build_manifold supports at most %d planes (got %d)", kMfMaxPlanes, P);
        return TSP_ERR_INVALID;
    }
    const size_t n = (size_t)P * R * C;
    if (n >= 0xFFFFFFFFull) {
        set_error("build_manifold: score volume too large (%zu voxels)", n);
        return TSP_ERR_INVALID;
    }
    unsigned long long* best = (unsigned long long*)d_scratch;
    TSP_CUDA(cudaMemsetAsync(best, 0, sizeof(unsigned long long), s));
    size_t blocks = (n + 255) / 256;
    if (blocks > (size_t)h->sm_count * 16) blocks = (size_t)h->sm_count * 16;
    argmax3d_kernel<<<(int)blocks, 256, 0, s>>>(d_score, n, best);
    TSP_LAUNCH_CHECK(h);
    const int ppad = ((P + 3) / 4 | 1) * 4;      // bytes per table: whole words, an odd number of them
    int chunk_cells = kMfTableBytes / (ppad + 1);
    const int longest = R > C ? R : C;
    if (chunk_cells > longest) chunk_cells = longest;
    chunk_cells = (chunk_cells + 15) / 16 * 16;
    const size_t smem = (size_t)chunk_cells * ppad + chunk_cells + (size_t)32 * ppad;
    {
        std::lock_guard<std::mutex> lock(h->mu);
        if (!h->manifold_attr) {
            TSP_CUDA(cudaFuncSetAttribute(manifold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            h->manifold_attr = true;
        }
    }
    manifold_kernel<<<1, kMfThreads, smem, s>>>(d_score, d_chosen, P, R, C, best, d_status, ppad, chunk_cells);
    TSP_LAUNCH_CHECK(h);
    return TSP_OK;
}

}  // namespace tsp
