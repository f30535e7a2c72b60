// Direct separable Gaussian FIR passes = scipy.ndimage.gaussian_filter(mode='nearest'), the only
// arithmetic primitive of the reference (basic_image_manipulations.py:373-390).  One kernel per
// axis; each pass stores its output dtype before the next pass like scipy does (float32 rounds,
// uint16 truncates).  Two accumulation policies:
//   fp32  - sequential FMA accumulation, register blocked (TSP_MODE_EXACT)
//   fp64  - scipy's own summation order: centre tap first, then the symmetric pairs from the
//           outermost inwards, (a + b) * w added with separate mul/add (TSP_MODE_BITEXACT)
#include <math.h>

#include <type_traits>

#include "common.cuh"

namespace tsp {

std::vector<double> gaussian_taps(double sigma) {
    if (sigma <= 1e-15) return std::vector<double>(1, 1.0);      // scipy skips such an axis: identity
    const int radius = (int)(4.0 * sigma + 0.5);
    std::vector<double> w(2 * radius + 1);
    double sum = 0.0;
    const double s2 = sigma * sigma;
    for (int k = -radius; k <= radius; ++k) {
        const double v = exp(-0.5 / s2 * (double)k * (double)k);
        w[k + radius] = v;
        sum += v;
    }
    for (auto& v : w) v /= sum;
    return w;
}

int get_taps(tsp_handle* h, double sigma, DeviceTaps* out) {
    char key[64];
    snprintf(key, sizeof key, "g%.17g", sigma);
    std::lock_guard<std::mutex> lock(h->mu);
    auto it = h->taps.find(key);
    if (it != h->taps.end()) {
        *out = it->second;
        return TSP_OK;
    }
    std::vector<double> w = gaussian_taps(sigma);
    const int n = (int)w.size();
    std::vector<float> w32(n + 2 * kTapPad, 0.0f);
    for (int i = 0; i < n; ++i) w32[i + kTapPad] = (float)w[i];
    double* d64 = nullptr;
    float* d32 = nullptr;
    TSP_CUDA(cudaMalloc(&d64, n * sizeof(double)));
    TSP_CUDA(cudaMalloc(&d32, w32.size() * sizeof(float)));
    TSP_CUDA(cudaMemcpy(d64, w.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    TSP_CUDA(cudaMemcpy(d32, w32.data(), w32.size() * sizeof(float), cudaMemcpyHostToDevice));
    DeviceTaps t;
    t.radius = (n - 1) / 2;
    t.w64 = d64;
    t.w32 = d32 + kTapPad;
    h->taps[key] = t;
    *out = t;
    return TSP_OK;
}

// ---- element I/O ------------------------------------------------------------------------------
__device__ __forceinline__ float load_as_float(const float* p) { return *p; }
__device__ __forceinline__ float load_as_float(const uint16_t* p) { return (float)*p; }
__device__ __forceinline__ void store_from(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_from(float* p, double v) { *p = (float)v; }          // round to nearest
__device__ __forceinline__ void store_from(uint16_t* p, double v) { *p = (uint16_t)__double2uint_rz(v); }
__device__ __forceinline__ void store_from(uint16_t* p, float v) { *p = (uint16_t)__float2uint_rz(v); }

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// scipy order on a symmetric kernel: tmp = x[0]*w[0]; for j=-r..-1: tmp += (x[j] + x[-j]) * w[j]
// `get(k)` returns the sample at offset k as float; wc points at the centre tap.
template <typename Get>
__device__ __forceinline__ double scipy_line_sum(Get get, const double* wc, int r) {
    double tmp = __dmul_rn((double)get(0), wc[0]);
    for (int j = -r; j < 0; ++j) {
        const double pair = __dadd_rn((double)get(j), (double)get(-j));
        tmp = __dadd_rn(tmp, __dmul_rn(pair, wc[j]));
    }
    return tmp;
}

// ---- axis 0 (z): one thread per (y,x) column position ----------------------------------------
template <typename T, bool FP64>
__global__ void fir_z_kernel(const T* __restrict__ in, T* __restrict__ out, int Z, size_t plane, int r,
                             const double* __restrict__ w64, const float* __restrict__ w32) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < plane; p += stride) {
        for (int z = 0; z < Z; ++z) {
            if (FP64) {
                auto get = [&](int k) { return load_as_float(in + (size_t)clampi(z + k, 0, Z - 1) * plane + p); };
                store_from(out + (size_t)z * plane + p, scipy_line_sum(get, w64 + r, r));
            } else {
                float acc = 0.f;
                for (int k = -r; k <= r; ++k)
                    acc = fmaf(w32[k + r], load_as_float(in + (size_t)clampi(z + k, 0, Z - 1) * plane + p), acc);
                store_from(out + (size_t)z * plane + p, acc);
            }
        }
    }
}

// ---- axis 1 (y): 32-column x TY-row tile with a +-r row halo in shared memory -------------------
constexpr int kFirYWarps = 8;
constexpr int kFirYOut = 16;                       // outputs per thread along y
constexpr int kFirYTile = kFirYWarps * kFirYOut;   // 128 rows

template <typename T, bool FP64>
__global__ void __launch_bounds__(32 * kFirYWarps)
fir_y_kernel(const T* __restrict__ in, T* __restrict__ out, int Y, int X, int r,
             const double* __restrict__ w64, const float* __restrict__ w32) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int rows = kFirYTile + 2 * r;
    const int rows_alloc = rows + kFirYOut;                       // slack so the blocked loop may overrun
    float* tile = reinterpret_cast<float*>(smem_raw);               // [rows_alloc][32]
    float* wsh = tile + (size_t)rows_alloc * 32;                    // fp32: padded taps
    double* wsh64 = reinterpret_cast<double*>(wsh);                 // fp64: taps (aliased, one policy per launch)

    const int lane = threadIdx.x, warp = threadIdx.y;
    const int x = blockIdx.x * 32 + lane;
    const int y0 = blockIdx.y * kFirYTile;
    const size_t zoff = (size_t)blockIdx.z * Y * X;
    const int xc = x < X ? x : X - 1;

    for (int i = warp; i < rows_alloc; i += kFirYWarps) {
        const int yy = clampi(y0 - r + i, 0, Y - 1);
        tile[i * 32 + lane] = load_as_float(in + zoff + (size_t)yy * X + xc);
    }
    const int tid = warp * 32 + lane;
    if (FP64) {
        for (int i = tid; i < 2 * r + 1; i += 32 * kFirYWarps) wsh64[i] = w64[i];
    } else {
        // wsh[i] = w[i] for i in [0, 2r], zero beyond (blocked loop overruns by < kFirYOut)
        for (int i = tid; i < 2 * r + 1 + 2 * kFirYOut; i += 32 * kFirYWarps) wsh[i] = i <= 2 * r ? w32[i] : 0.f;
    }
    __syncthreads();

    const int ybase = warp * kFirYOut;            // first output row of this thread inside the tile
    if (FP64) {
        for (int j = 0; j < kFirYOut; ++j) {
            const int y = y0 + ybase + j;
            if (y >= Y || x >= X) break;
            const float* centre = tile + (size_t)(ybase + j + r) * 32 + lane;
            auto get = [&](int k) { return centre[k * 32]; };
            store_from(out + zoff + (size_t)y * X + x, scipy_line_sum(get, wsh64 + r, r));
        }
    } else {
        float acc[kFirYOut];
        float wc[kFirYOut];
#pragma unroll
        for (int j = 0; j < kFirYOut; ++j) { acc[j] = 0.f; wc[j] = 0.f; }
        const float* col = tile + (size_t)ybase * 32 + lane;
        // input ii (relative to ybase) feeds output j with tap w[ii - j]; circular tap window wc[]
        for (int ii = 0; ii < kFirYOut + 2 * r; ii += kFirYOut) {
#pragma unroll
            for (int u = 0; u < kFirYOut; ++u) {
                wc[u] = wsh[ii + u];
                const float v = col[(size_t)(ii + u) * 32];
#pragma unroll
                for (int j = 0; j < kFirYOut; ++j) acc[j] = fmaf(wc[(u - j + kFirYOut) % kFirYOut], v, acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < kFirYOut; ++j) {
            const int y = y0 + ybase + j;
            if (y < Y && x < X) store_from(out + zoff + (size_t)y * X + x, acc[j]);
        }
    }
}

// ---- axis 2 (x): kFirXRows rows x kFirXTile outputs, +-r column halo in shared memory ----------
constexpr int kFirXTile = 256;
constexpr int kFirXRows = 4;
constexpr int kFirXThreads = 256;    // 64 threads (x4 outputs) per row

template <typename T, bool FP64>
__global__ void __launch_bounds__(kFirXThreads)
fir_x_kernel(const T* __restrict__ in, T* __restrict__ out, size_t total_rows, int X, int r,
             const double* __restrict__ w64, const float* __restrict__ w32) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int row_len = ((kFirXTile + 2 * r + 8) + 3) & ~3;
    float* tile = reinterpret_cast<float*>(smem_raw);                // [kFirXRows][row_len]
    float* wsh = tile + (size_t)kFirXRows * row_len;                  // fp32: 4 zeros | taps | zeros
    double* wsh64 = reinterpret_cast<double*>(wsh);

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kFirXTile;
    const size_t group = (size_t)blockIdx.z * gridDim.y + blockIdx.y;
    const size_t row0 = group * kFirXRows;                           // row index over Z*Y
    if (row0 >= total_rows) return;

    for (int rr = 0; rr < kFirXRows; ++rr) {
        size_t row = row0 + rr;
        if (row >= total_rows) row = total_rows - 1;
        const T* src = in + row * X;
        for (int i = tid; i < row_len; i += kFirXThreads)
            tile[rr * row_len + i] = load_as_float(src + clampi(x0 - r + i, 0, X - 1));
    }
    if (FP64) {
        for (int i = tid; i < 2 * r + 1; i += kFirXThreads) wsh64[i] = w64[i];
    } else {
        for (int i = tid; i < 2 * r + 1 + 16; i += kFirXThreads) {
            const int k = i - 4;
            wsh[i] = (k >= 0 && k <= 2 * r) ? w32[k] : 0.f;
        }
    }
    __syncthreads();

    if (FP64) {
        for (int o = tid; o < kFirXRows * kFirXTile; o += kFirXThreads) {
            const int rr = o / kFirXTile, xo = o % kFirXTile;
            const size_t row = row0 + rr;
            const int x = x0 + xo;
            if (row >= total_rows || x >= X) continue;
            const float* centre = tile + rr * row_len + xo + r;
            auto get = [&](int k) { return centre[k]; };
            store_from(out + row * X + x, scipy_line_sum(get, wsh64 + r, r));
        }
    } else {
        const int rr = tid / 64, t = tid % 64;
        const size_t row = row0 + rr;
        const float4* src4 = reinterpret_cast<const float4*>(tile + rr * row_len + 4 * t);
        const float4* w4 = reinterpret_cast<const float4*>(wsh);      // w4[c] = taps 4c-4 .. 4c-1
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        float4 wp = w4[0];                                            // taps -4..-1 (zeros)
        const int nchunks = (2 * r + 4 + 3) / 4;
        for (int c = 0; c < nchunks; ++c) {
            const float4 v = src4[c];
            const float4 wn = w4[c + 1];                              // taps 4c .. 4c+3
            // input m (0..3) of this chunk is sample 4c+m; output j uses tap 4c+m-j
            const float wv[7] = {wp.y, wp.z, wp.w, wn.x, wn.y, wn.z, wn.w};   // taps 4c-3 .. 4c+3
            const float in4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j] = fmaf(wv[m - j + 3], in4[m], acc[j]);
            wp = wn;
        }
        if (row < total_rows) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int x = x0 + 4 * t + j;
                if (x < X) store_from(out + row * X + x, acc[j]);
            }
        }
    }
}

// ---- fp32 passes at packed rate (TSP_MODE_EXACT, the general path, the materialised band mask) ------------------
// On sm_100 a scalar FFMA issues every second cycle per scheduler; the packed FFMA2 (two FMAs per lane) issues at
// the same rate, so only packed arithmetic reaches the FP32 peak.  Both in-plane passes therefore run the same
// register-blocked line filter on PAIRS of lines: a thread owns 16 consecutive outputs of two neighbouring lines
// (float2 accumulators), the taps stream through a 16-entry circular window of (w, w) pairs, every input pair costs
// two 8-byte shared loads and 16 FFMA2.  241 taps (sigma = 30): 256 FFMA2 per 16 x 2 outputs, 94 % of them useful.
//   y pass: lines = image columns; a lane owns a column pair (adjacent floats: natural float2 loads, conflict free)
//   x pass: lines = image rows; a lane owns a row pair, interleaved as float2 in shared memory with an odd pitch
//           (conflict-free 8-byte loads down the lanes); the rows are loaded coalesced and transposed on the way in
constexpr int kL2Out = 16;                 // outputs per thread along the line
constexpr int kL2Warps = 8;                // warps per CTA = segments of 16 outputs along the line
constexpr int kL2Tile = kL2Warps * kL2Out; // 128 outputs along the line per CTA

__device__ __forceinline__ void line_fir16(const float2* __restrict__ col, int stride, const float2* __restrict__ wsh2,
                                           int r, float2 (&acc)[kL2Out]) {
    float2 wc[kL2Out];
#pragma unroll
    for (int j = 0; j < kL2Out; ++j) {
        acc[j] = make_float2(0.f, 0.f);
        wc[j] = make_float2(0.f, 0.f);
    }
    // input ii (relative to the first output) feeds output j with tap ii - j; wsh2[i] = tap i for i in [0, 2r], 0 beyond
    for (int ii = 0; ii < kL2Out + 2 * r; ii += kL2Out) {
#pragma unroll
        for (int u = 0; u < kL2Out; ++u) {
            wc[u] = wsh2[ii + u];
            const float2 v = col[(size_t)(ii + u) * stride];
#pragma unroll
            for (int j = 0; j < kL2Out; ++j) acc[j] = __ffma2_rn(v, wc[(u - j + kL2Out) % kL2Out], acc[j]);
        }
    }
}

// y pass: CTA = 64 columns (32 pairs) x 128 output rows of one plane
__global__ void __launch_bounds__(32 * kL2Warps)
fir_y2_kernel(const float* __restrict__ in, float* __restrict__ out, int Y, int X, int r, const float* __restrict__ w32) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int rows_alloc = kL2Tile + 2 * r + kL2Out;                 // slack: the blocked loop overruns by < 16 inputs
    float* tile = reinterpret_cast<float*>(smem_raw);               // [rows_alloc][64]
    float2* wsh2 = reinterpret_cast<float2*>(tile + (size_t)rows_alloc * 64);
    const int lane = threadIdx.x, warp = threadIdx.y, tid = warp * 32 + lane;
    const int x0 = blockIdx.x * 64, y0 = blockIdx.y * kL2Tile;
    const size_t zoff = (size_t)blockIdx.z * Y * X;
    if (x0 + 64 <= X && (X & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {      // 16-byte loads
        float4* tile4 = reinterpret_cast<float4*>(tile);
        for (int i = tid; i < rows_alloc * 16; i += 32 * kL2Warps) {
            const int yy = clampi(y0 - r + (i >> 4), 0, Y - 1);
            tile4[i] = __ldg(reinterpret_cast<const float4*>(in + zoff + (size_t)yy * X + x0) + (i & 15));
        }
    } else {
        for (int i = tid; i < rows_alloc * 64; i += 32 * kL2Warps) {
            const int yy = clampi(y0 - r + i / 64, 0, Y - 1), xx = min(x0 + (i & 63), X - 1);
            tile[i] = __ldg(in + zoff + (size_t)yy * X + xx);
        }
    }
    for (int i = tid; i < 2 * r + 1 + 2 * kL2Out; i += 32 * kL2Warps) {
        const float w = i <= 2 * r ? w32[i] : 0.f;
        wsh2[i] = make_float2(w, w);
    }
    __syncthreads();
    float2 acc[kL2Out];
    line_fir16(reinterpret_cast<const float2*>(tile) + (size_t)(warp * kL2Out) * 32 + lane, 32, wsh2, r, acc);
    const int x = x0 + 2 * lane;
#pragma unroll
    for (int j = 0; j < kL2Out; ++j) {
        const int y = y0 + warp * kL2Out + j;
        if (y >= Y || x >= X) continue;
        float* dst = out + zoff + (size_t)y * X + x;
        if (x + 1 < X && (X & 1) == 0) *reinterpret_cast<float2*>(dst) = acc[j];
        else {
            dst[0] = acc[j].x;
            if (x + 1 < X) dst[1] = acc[j].y;
        }
    }
}

// x pass: CTA = 64 rows (32 pairs) x 128 output columns; rows run over the flattened (Z*Y) row index
__global__ void __launch_bounds__(32 * kL2Warps)
fir_x2_kernel(const float* __restrict__ in, float* __restrict__ out, size_t total_rows, int X, int r,
              const float* __restrict__ w32) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int xlen = kL2Tile + 2 * r + kL2Out;
    const int pitch = xlen | 1;                                      // odd, in float2 units
    float2* tile2 = reinterpret_cast<float2*>(smem_raw);            // [32 row pairs][pitch]
    float2* wsh2 = tile2 + (size_t)32 * pitch;
    const int lane = threadIdx.x, warp = threadIdx.y, tid = warp * 32 + lane;
    const int x0 = blockIdx.x * kL2Tile;
    const size_t row0 = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * 64;
    if (row0 >= total_rows) return;
    float* tile = reinterpret_cast<float*>(tile2);
    for (int rr = warp; rr < 64; rr += kL2Warps) {                   // one warp per row: coalesced along x
        size_t row = row0 + rr;
        if (row >= total_rows) row = total_rows - 1;
        const float* src = in + row * X;
        float* dst = tile + ((size_t)(rr >> 1) * pitch) * 2 + (rr & 1);
        if ((r & 3) == 0 && (X & 3) == 0 && x0 - r >= 0 && x0 - r + xlen <= X &&
            (reinterpret_cast<uintptr_t>(in) & 15) == 0) {                                     // 16-byte loads
            const float4* src4 = reinterpret_cast<const float4*>(src + (x0 - r));
            for (int i = lane; i < xlen / 4; i += 32) {
                const float4 v = __ldg(src4 + i);
                dst[8 * i] = v.x; dst[8 * i + 2] = v.y; dst[8 * i + 4] = v.z; dst[8 * i + 6] = v.w;
            }
        } else {
            for (int i = lane; i < xlen; i += 32) dst[2 * i] = __ldg(src + clampi(x0 - r + i, 0, X - 1));
        }
    }
    for (int i = tid; i < 2 * r + 1 + 2 * kL2Out; i += 32 * kL2Warps) {
        const float w = i <= 2 * r ? w32[i] : 0.f;
        wsh2[i] = make_float2(w, w);
    }
    __syncthreads();
    float2 acc[kL2Out];
    line_fir16(tile2 + (size_t)lane * pitch + warp * kL2Out, 1, wsh2, r, acc);
    const int xo = x0 + warp * kL2Out;
    const size_t ra = row0 + 2 * lane, rb = ra + 1;
    const bool vec = (X & 3) == 0 && xo + kL2Out <= X;
    if (ra < total_rows) {
        float* da = out + ra * X + xo;
        if (vec) {
#pragma unroll
            for (int q = 0; q < kL2Out / 4; ++q)
                reinterpret_cast<float4*>(da)[q] = make_float4(acc[4 * q].x, acc[4 * q + 1].x, acc[4 * q + 2].x, acc[4 * q + 3].x);
        } else {
#pragma unroll
            for (int j = 0; j < kL2Out; ++j)
                if (xo + j < X) da[j] = acc[j].x;
        }
    }
    if (rb < total_rows) {
        float* db = out + rb * X + xo;
        if (vec) {
#pragma unroll
            for (int q = 0; q < kL2Out / 4; ++q)
                reinterpret_cast<float4*>(db)[q] = make_float4(acc[4 * q].y, acc[4 * q + 1].y, acc[4 * q + 2].y, acc[4 * q + 3].y);
        } else {
#pragma unroll
            for (int j = 0; j < kL2Out; ++j)
                if (xo + j < X) db[j] = acc[j].y;
        }
    }
}

// Short in-plane filters (sigma = 1 and 2 of SP:37 / SP:70-71: radius <= 8) as ONE kernel: a 32 x 128 output tile with
// its halo goes to shared memory, the y pass writes an intermediate tile (float32, the rounding point between scipy's
// passes), the x pass reads it - one global read and one write instead of two of each.  Taps are zero padded to the
// compile-time radius R; sums run over the taps in ascending order like the unfused passes (adding w = 0 terms is exact).
constexpr int kYXTY = 32, kYXTX = 128;

template <int R>
__global__ void __launch_bounds__(256) fir_yx_fused_kernel(const float* __restrict__ in, float* __restrict__ out, int Y,
                                                           int X, int ry, int rx, const float* __restrict__ wy32,
                                                           const float* __restrict__ wx32) {
    constexpr int IW = kYXTX + 2 * R, IH = kYXTY + 2 * R, NT = 2 * R + 1;
    __shared__ float in_s[IH][IW];
    __shared__ float mid_s[kYXTY][IW];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * kYXTX, y0 = blockIdx.y * kYXTY;
    const size_t zoff = (size_t)blockIdx.z * Y * X;
    float wy[NT], wx[NT];                          // tap k multiplies the sample at offset k - R
#pragma unroll
    for (int k = 0; k < NT; ++k) {
        const int ky = k - R + ry, kx = k - R + rx;
        wy[k] = (ky >= 0 && ky <= 2 * ry) ? __ldg(wy32 + ky) : 0.f;
        wx[k] = (kx >= 0 && kx <= 2 * rx) ? __ldg(wx32 + kx) : 0.f;
    }
    for (int i = tid; i < IH * IW; i += 256) {
        const int r = i / IW, c = i - r * IW;
        const int yy = clampi(y0 - R + r, 0, Y - 1), xx = clampi(x0 - R + c, 0, X - 1);
        in_s[r][c] = __ldg(in + zoff + (size_t)yy * X + xx);
    }
    __syncthreads();
    // y pass: task = (column, group of 8 output rows)
    for (int task = tid; task < IW * (kYXTY / 8); task += 256) {
        const int c = task % IW, g = task / IW;
        float v[8 + 2 * R];
#pragma unroll
        for (int i = 0; i < 8 + 2 * R; ++i) v[i] = in_s[8 * g + i][c];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < NT; ++k) a = fmaf(wy[k], v[j + k], a);
            mid_s[8 * g + j][c] = a;
        }
    }
    __syncthreads();
    // x pass: task = (row, group of 8 output columns)
    for (int task = tid; task < kYXTY * (kYXTX / 8); task += 256) {
        const int g = task % (kYXTX / 8), r = task / (kYXTX / 8);
        const int y = y0 + r, x = x0 + 8 * g;
        if (y >= Y || x >= X) continue;
        float v[8 + 2 * R];
#pragma unroll
        for (int i = 0; i < 8 + 2 * R; ++i) v[i] = mid_s[r][8 * g + i];
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < NT; ++k) a = fmaf(wx[k], v[j + k], a);
            o[j] = a;
        }
        float* dst = out + zoff + (size_t)y * X + x;
        if (x + 7 < X && (X & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
            reinterpret_cast<float4*>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
            reinterpret_cast<float4*>(dst)[1] = make_float4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (x + j < X) dst[j] = o[j];
        }
    }
}

// both in-plane passes of a float32 blur, d_in -> d_out (d_in is not modified); false when the radii are too large
static bool fir_yx_fused(tsp_handle* h, const float* d_in, float* d_out, int Z, int Y, int X, const DeviceTaps& ty,
                         const DeviceTaps& tx, cudaStream_t s) {
    const int r = ty.radius > tx.radius ? ty.radius : tx.radius;
    if (r > 8) return false;
    dim3 grid((X + kYXTX - 1) / kYXTX, (Y + kYXTY - 1) / kYXTY, Z);
    if (r <= 4) fir_yx_fused_kernel<4><<<grid, 256, 0, s>>>(d_in, d_out, Y, X, ty.radius, tx.radius, ty.w32, tx.w32);
    else fir_yx_fused_kernel<8><<<grid, 256, 0, s>>>(d_in, d_out, Y, X, ty.radius, tx.radius, ty.w32, tx.w32);
    h->launches++;
    return true;
}

// ---- streaming form of the long line filters (radius 9 .. 120: the sigma = 30 passes) ----------------------------
// The tiled kernels above re-read a 2r-row halo per 128 outputs (3 x the volume at r = 120) and load / compute in
// turns.  Here a CTA owns its lines for their whole length: a shared-memory RING of 384 line positions = exactly the
// inputs of one step of 128 outputs at r = 120 (96 KB: TWO CTAs per SM).  Every input is fetched once (cp.async);
// the 128 positions of the next step are requested as soon as every warp has finished the current step - they
// overwrite the 128 oldest - and arrive while the CTA stores its outputs and the SM's other CTA computes: with one
// 128 KB ring (512 positions, loads behind the CTA's own FMAs) the SM held 8 warps and the FMA pipe idled 24 % of
// the time (ncu: sm__pipe_fma_cycles_active 76 %, occupancy 12.5 %).  Ring coordinate rho = position + r, so the
// first input of an output block sits at a multiple of 16: the 16-input blocks of the register filter never
// straddle the ring's wrap-around (384 = 24 x 16).
constexpr int kRingLen = 384;                         // line positions in the ring (a multiple of 16)
constexpr int kStreamMaxR = (kRingLen - kL2Tile - kL2Out) / 2;            // 120
__device__ __forceinline__ int ring_slot(int rho) { return rho % kRingLen; }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4s(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

// the register filter of line_fir16 reading a ring: block b (16 inputs) starts at ring position (rho0 + 16 b) & 511
template <typename Index>
__device__ __forceinline__ void line_fir16_ring(Index at, int rho0, const float2* __restrict__ wsh2, int r,
                                                float2 (&acc)[kL2Out]) {
    float2 wc[kL2Out];
#pragma unroll
    for (int j = 0; j < kL2Out; ++j) {
        acc[j] = make_float2(0.f, 0.f);
        wc[j] = make_float2(0.f, 0.f);
    }
    int pos = ring_slot(rho0);
    for (int ii = 0; ii < kL2Out + 2 * r; ii += kL2Out) {
        const float2* blk = at(pos);
        pos = pos + kL2Out >= kRingLen ? pos + kL2Out - kRingLen : pos + kL2Out;
#pragma unroll
        for (int u = 0; u < kL2Out; ++u) {
            wc[u] = wsh2[ii + u];
            const float2 v = at.step(blk, u);
#pragma unroll
            for (int j = 0; j < kL2Out; ++j) acc[j] = __ffma2_rn(v, wc[(u - j + kL2Out) % kL2Out], acc[j]);
        }
    }
}

struct YRingIndex {          // ring[rho][32 column pairs]: a lane owns a column pair
    const float2* base;      // + lane
    __device__ __forceinline__ const float2* operator()(int rho) const { return base + (size_t)rho * 32; }
    __device__ __forceinline__ float2 step(const float2* blk, int u) const { return blk[u * 32]; }
};
struct XRingIndex {          // ring[32 row pairs][pitch]: a lane owns a row pair
    const float2* base;      // + lane * pitch
    __device__ __forceinline__ const float2* operator()(int rho) const { return base + rho; }
    __device__ __forceinline__ float2 step(const float2* blk, int u) const { return blk[u]; }
};

// y pass: CTA = 64 columns of one plane, all rows
__global__ void __launch_bounds__(32 * kL2Warps, 2)
fir_y_stream_kernel(const float* __restrict__ in, float* __restrict__ out, int Y, int X, int r,
                    const float* __restrict__ w32) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);                           // [512][64]
    float2* wsh2 = reinterpret_cast<float2*>(ring + (size_t)kRingLen * 64);
    const int lane = threadIdx.x, warp = threadIdx.y, tid = warp * 32 + lane;
    const int x0 = blockIdx.x * 64;
    const size_t zoff = (size_t)blockIdx.z * Y * X;
    for (int i = tid; i < 2 * r + 1 + 2 * kL2Out; i += 32 * kL2Warps) {
        const float w = i <= 2 * r ? w32[i] : 0.f;
        wsh2[i] = make_float2(w, w);
    }
    // rows rho in [lo, hi) of the ring <- image rows clamp(rho - r): 16 float4 per row (x0 + 64 <= X, X % 4 == 0)
    auto fetch = [&](int lo, int hi) {
        for (int i = tid; i < (hi - lo) * 16; i += 32 * kL2Warps) {
            const int rho = lo + (i >> 4);
            const int yy = clampi(rho - r, 0, Y - 1);
            cp_async16(ring + (size_t)ring_slot(rho) * 64 + 4 * (i & 15), in + zoff + (size_t)yy * X + x0 + 4 * (i & 15));
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int need = kL2Tile + 2 * r + kL2Out;                                  // inputs of one step (<= kRingLen)
    fetch(0, need);
    const int nsteps = (Y + kL2Tile - 1) / kL2Tile;
    for (int s = 0; s < nsteps; ++s) {
        const int y0 = s * kL2Tile;
        asm volatile("cp.async.wait_group 0;" ::: "memory");                    // this step's rows have landed
        __syncthreads();
        float2 acc[kL2Out];
        const YRingIndex at{reinterpret_cast<const float2*>(ring) + lane};
        line_fir16_ring(at, y0 + warp * kL2Out, wsh2, r, acc);
        __syncthreads();                                  // everyone is done with the rows the next fetch overwrites
        if (s + 1 < nsteps) fetch(y0 + need, y0 + need + kL2Tile);
        const int x = x0 + 2 * lane;
#pragma unroll
        for (int j = 0; j < kL2Out; ++j) {
            const int y = y0 + warp * kL2Out + j;
            if (y < Y) *reinterpret_cast<float2*>(out + zoff + (size_t)y * X + x) = acc[j];
        }
    }
}

// x pass: CTA = 64 rows (32 pairs) of the flattened (Z*Y) row index, all columns
__global__ void __launch_bounds__(32 * kL2Warps, 2)
fir_x_stream_kernel(const float* __restrict__ in, float* __restrict__ out, size_t total_rows, int X, int r,
                    const float* __restrict__ w32) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int pitch = kRingLen + 1;                                         // odd, in float2 units
    float2* ring2 = reinterpret_cast<float2*>(smem_raw);                        // [32 row pairs][pitch]
    float2* wsh2 = ring2 + (size_t)32 * pitch;
    float* ringf = reinterpret_cast<float*>(ring2);
    const int lane = threadIdx.x, warp = threadIdx.y, tid = warp * 32 + lane;
    const size_t row0 = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 64;
    if (row0 >= total_rows) return;
    for (int i = tid; i < 2 * r + 1 + 2 * kL2Out; i += 32 * kL2Warps) {
        const float w = i <= 2 * r ? w32[i] : 0.f;
        wsh2[i] = make_float2(w, w);
    }
    // columns rho in [lo, hi) <- image columns clamp(rho - r), every row of the CTA: a warp takes rows warp, warp+8, ..
    auto fetch = [&](int lo, int hi) {
        for (int rr = warp; rr < 64; rr += kL2Warps) {
            size_t row = row0 + rr;
            if (row >= total_rows) row = total_rows - 1;
            const float* src = in + row * X;
            float* dst = ringf + ((size_t)(rr >> 1) * pitch) * 2 + (rr & 1);
            for (int rho = lo + lane; rho < hi; rho += 32)
                cp_async4s(dst + 2 * ring_slot(rho), src + clampi(rho - r, 0, X - 1));
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int need = kL2Tile + 2 * r + kL2Out;
    fetch(0, need);
    const int nsteps = (X + kL2Tile - 1) / kL2Tile;
    const size_t ra = row0 + 2 * lane, rb = ra + 1;
    for (int s = 0; s < nsteps; ++s) {
        const int xs = s * kL2Tile;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        float2 acc[kL2Out];
        const XRingIndex at{ring2 + (size_t)lane * pitch};
        line_fir16_ring(at, xs + warp * kL2Out, wsh2, r, acc);
        __syncthreads();                                  // everyone is done with the columns the next fetch overwrites
        if (s + 1 < nsteps) fetch(xs + need, xs + need + kL2Tile);
        const int xo = xs + warp * kL2Out;
        const bool vec = (X & 3) == 0 && xo + kL2Out <= X;
        if (ra < total_rows) {
            float* da = out + ra * X + xo;
            if (vec) {
#pragma unroll
                for (int q = 0; q < kL2Out / 4; ++q)
                    reinterpret_cast<float4*>(da)[q] = make_float4(acc[4 * q].x, acc[4 * q + 1].x, acc[4 * q + 2].x, acc[4 * q + 3].x);
            } else {
#pragma unroll
                for (int j = 0; j < kL2Out; ++j)
                    if (xo + j < X) da[j] = acc[j].x;
            }
        }
        if (rb < total_rows) {
            float* db = out + rb * X + xo;
            if (vec) {
#pragma unroll
                for (int q = 0; q < kL2Out / 4; ++q)
                    reinterpret_cast<float4*>(db)[q] = make_float4(acc[4 * q].y, acc[4 * q + 1].y, acc[4 * q + 2].y, acc[4 * q + 3].y);
            } else {
#pragma unroll
                for (int j = 0; j < kL2Out; ++j)
                    if (xo + j < X) db[j] = acc[j].y;
            }
        }
    }
}

// z pass: a thread marches one group of four columns through the planes with the 2R+1 inputs it needs in registers:
// every voxel is read once and written once (the generic kernel re-reads each input 2R+1 times through the caches)
template <int R>
__global__ void __launch_bounds__(256) fir_z_march_kernel(const float* __restrict__ in, float* __restrict__ out, int Z,
                                                          size_t plane4, const float* __restrict__ w32) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= plane4) return;
    const float4* src = reinterpret_cast<const float4*>(in) + p;
    float4* dst = reinterpret_cast<float4*>(out) + p;
    float w[2 * R + 1];
#pragma unroll
    for (int k = 0; k <= 2 * R; ++k) w[k] = __ldg(w32 + k);
    float4 win[2 * R + 1];                        // win[k] = plane z - R + k (edge replicated)
#pragma unroll
    for (int k = 0; k <= 2 * R; ++k) win[k] = __ldg(src + (size_t)clampi(k - R, 0, Z - 1) * plane4);
    for (int z0 = 0; z0 < Z; z0 += 2 * R + 1) {
#pragma unroll
        for (int u = 0; u <= 2 * R; ++u) {        // the window rotates through its 2R+1 names: no register moves
            const int z = z0 + u;
            if (z >= Z) break;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k <= 2 * R; ++k) {
                const float4 v = win[(u + k) % (2 * R + 1)];
                a.x = fmaf(w[k], v.x, a.x); a.y = fmaf(w[k], v.y, a.y);
                a.z = fmaf(w[k], v.z, a.z); a.w = fmaf(w[k], v.w, a.w);
            }
            dst[(size_t)z * plane4] = a;
            win[u] = __ldg(src + (size_t)clampi(z + R + 1, 0, Z - 1) * plane4);   // slot of plane z - R becomes z + R + 1
        }
    }
}

template <typename T>
int launch_fir_axis(tsp_handle* h, const T* d_in, T* d_out, int Z, int Y, int X, int axis,
                    const DeviceTaps& taps, bool fp64, cudaStream_t s) {
    const int r = taps.radius;
    if constexpr (std::is_same<T, float>::value) {
        if (!fp64) {
            const float* fin = d_in;
            float* fout = d_out;
            const size_t plane = (size_t)Y * X;
            const bool al16 = ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15) == 0;
            if (axis == 0 && (r == 2 || r == 4 || r == 1 || r == 0) && plane % 4 == 0 && al16) {
                const size_t plane4 = plane / 4;
                const unsigned blocks = (unsigned)((plane4 + 255) / 256);
                if (r == 0) fir_z_march_kernel<0><<<blocks, 256, 0, s>>>(fin, fout, Z, plane4, taps.w32);
                else if (r == 1) fir_z_march_kernel<1><<<blocks, 256, 0, s>>>(fin, fout, Z, plane4, taps.w32);
                else if (r == 2) fir_z_march_kernel<2><<<blocks, 256, 0, s>>>(fin, fout, Z, plane4, taps.w32);
                else fir_z_march_kernel<4><<<blocks, 256, 0, s>>>(fin, fout, Z, plane4, taps.w32);
                TSP_LAUNCH_CHECK(h);
                return TSP_OK;
            }
            if (axis == 1 && r > 8 && r <= kStreamMaxR && X % 64 == 0 && al16 && Y >= kL2Tile) {
                const size_t smem = (size_t)kRingLen * 64 * sizeof(float) + (size_t)(2 * r + 1 + 2 * kL2Out) * sizeof(float2);
                TSP_CUDA(cudaFuncSetAttribute(fir_y_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                dim3 grid(X / 64, 1, Z);
                fir_y_stream_kernel<<<grid, dim3(32, kL2Warps), smem, s>>>(fin, fout, Y, X, r, taps.w32);
                TSP_LAUNCH_CHECK(h);
                return TSP_OK;
            }
            if (axis == 2 && r > 8 && r <= kStreamMaxR && X >= kL2Tile) {
                const size_t smem = (size_t)32 * (kRingLen + 1) * sizeof(float2) + (size_t)(2 * r + 1 + 2 * kL2Out) * sizeof(float2);
                TSP_CUDA(cudaFuncSetAttribute(fir_x_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                const size_t total_rows = (size_t)Z * Y;
                const size_t groups = (total_rows + 63) / 64;
                dim3 grid((unsigned)(groups < 32768 ? groups : 32768), 1, 1);
                grid.y = (unsigned)((groups + grid.x - 1) / grid.x);
                fir_x_stream_kernel<<<grid, dim3(32, kL2Warps), smem, s>>>(fin, fout, total_rows, X, r, taps.w32);
                TSP_LAUNCH_CHECK(h);
                return TSP_OK;
            }
            if (axis == 1) {
                const size_t smem = (size_t)(kL2Tile + 2 * r + kL2Out) * 64 * sizeof(float) +
                                    (size_t)(2 * r + 1 + 2 * kL2Out) * sizeof(float2);
                if (smem <= 200 * 1024) {
                    TSP_CUDA(cudaFuncSetAttribute(fir_y2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    dim3 grid((X + 63) / 64, (Y + kL2Tile - 1) / kL2Tile, Z);
                    fir_y2_kernel<<<grid, dim3(32, kL2Warps), smem, s>>>(fin, fout, Y, X, r, taps.w32);
                    TSP_LAUNCH_CHECK(h);
                    return TSP_OK;
                }
            }
            if (axis == 2) {
                const int xlen = kL2Tile + 2 * r + kL2Out;
                const size_t smem = (size_t)32 * (xlen | 1) * sizeof(float2) + (size_t)(2 * r + 1 + 2 * kL2Out) * sizeof(float2);
                if (smem <= 200 * 1024) {
                    TSP_CUDA(cudaFuncSetAttribute(fir_x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    const size_t total_rows = (size_t)Z * Y;
                    const size_t groups = (total_rows + 63) / 64;
                    dim3 grid((X + kL2Tile - 1) / kL2Tile, 1, 1);
                    grid.y = (unsigned)(groups < 32768 ? groups : 32768);
                    grid.z = (unsigned)((groups + grid.y - 1) / grid.y);
                    fir_x2_kernel<<<grid, dim3(32, kL2Warps), smem, s>>>(fin, fout, total_rows, X, r, taps.w32);
                    TSP_LAUNCH_CHECK(h);
                    return TSP_OK;
                }
            }
        }
    }
    if (axis == 0) {
        const size_t plane = (size_t)Y * X;
        const int threads = 256;
        size_t blocks = (plane + threads - 1) / threads;
        if (blocks > (size_t)h->sm_count * 32) blocks = (size_t)h->sm_count * 32;
        if (fp64)
            fir_z_kernel<T, true><<<(int)blocks, threads, 0, s>>>(d_in, d_out, Z, plane, r, taps.w64, taps.w32);
        else
            fir_z_kernel<T, false><<<(int)blocks, threads, 0, s>>>(d_in, d_out, Z, plane, r, taps.w64, taps.w32);
        TSP_LAUNCH_CHECK(h);
    } else if (axis == 1) {
        dim3 block(32, kFirYWarps);
        dim3 grid((X + 31) / 32, (Y + kFirYTile - 1) / kFirYTile, Z);
        const size_t rows_alloc = kFirYTile + 2 * r + kFirYOut;
        const size_t smem = rows_alloc * 32 * sizeof(float) + (size_t)(2 * r + 1 + 2 * kFirYOut) * sizeof(double);
        if (smem > 200 * 1024) {
            set_error("gaussian radius %d too large for the y pass", r);
            return TSP_ERR_INVALID;
        }
        if (fp64) {
            TSP_CUDA(cudaFuncSetAttribute(fir_y_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            fir_y_kernel<T, true><<<grid, block, smem, s>>>(d_in, d_out, Y, X, r, taps.w64, taps.w32);
        } else {
            TSP_CUDA(cudaFuncSetAttribute(fir_y_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            fir_y_kernel<T, false><<<grid, block, smem, s>>>(d_in, d_out, Y, X, r, taps.w64, taps.w32);
        }
        TSP_LAUNCH_CHECK(h);
    } else {
        const size_t total_rows = (size_t)Z * Y;
        const int row_len = ((kFirXTile + 2 * r + 8) + 3) & ~3;
        const size_t smem = (size_t)kFirXRows * row_len * sizeof(float) + (size_t)(2 * r + 1 + 16) * sizeof(double);
        if (smem > 200 * 1024) {
            set_error("gaussian radius %d too large for the x pass", r);
            return TSP_ERR_INVALID;
        }
        const size_t groups = (total_rows + kFirXRows - 1) / kFirXRows;
        dim3 grid((X + kFirXTile - 1) / kFirXTile, 1, 1);
        grid.y = (unsigned)(groups < 32768 ? groups : 32768);
        grid.z = (unsigned)((groups + grid.y - 1) / grid.y);
        if (fp64) {
            TSP_CUDA(cudaFuncSetAttribute(fir_x_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            fir_x_kernel<T, true><<<grid, kFirXThreads, smem, s>>>(d_in, d_out, total_rows, X, r, taps.w64, taps.w32);
        } else {
            TSP_CUDA(cudaFuncSetAttribute(fir_x_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            fir_x_kernel<T, false><<<grid, kFirXThreads, smem, s>>>(d_in, d_out, total_rows, X, r, taps.w64, taps.w32);
        }
        TSP_LAUNCH_CHECK(h);
    }
    return TSP_OK;
}

template <typename T>
int gaussian_blur(tsp_handle* h, const T* d_in, T* d_out, T* d_tmp, int Z, int Y, int X,
                  const double sigma[3], bool fp64, cudaStream_t s) {
    // three passes in -> out ; out -> tmp ; tmp -> out  (never writes d_in)
    DeviceTaps t[3];
    for (int a = 0; a < 3; ++a) {
        int rc = get_taps(h, sigma[a], &t[a]);
        if (rc) return rc;
    }
    if constexpr (std::is_same<T, float>::value) {
        // short in-plane filters: y and x in one kernel (z pass into d_tmp first).  Callers may pass d_tmp == d_in:
        // the marching z pass works in place (it never re-reads a plane it has written), the generic one does not
        const int rz = t[0].radius;
        const bool march = (rz == 0 || rz == 1 || rz == 2 || rz == 4) && ((size_t)Y * X) % 4 == 0 &&
                           ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_tmp)) & 15) == 0;
        if (!fp64 && t[1].radius <= 8 && t[2].radius <= 8 && (march || (const void*)d_tmp != (const void*)d_in)) {
            int rc = launch_fir_axis<T>(h, d_in, d_tmp, Z, Y, X, 0, t[0], false, s);
            if (rc) return rc;
            fir_yx_fused(h, d_tmp, d_out, Z, Y, X, t[1], t[2], s);
            if (cudaGetLastError() != cudaSuccess) {
                set_error("fused y/x pass failed to launch");
                return TSP_ERR_CUDA;
            }
            return TSP_OK;
        }
    }
    int rc = launch_fir_axis<T>(h, d_in, d_out, Z, Y, X, 0, t[0], fp64, s);
    if (rc) return rc;
    rc = launch_fir_axis<T>(h, d_out, d_tmp, Z, Y, X, 1, t[1], fp64, s);
    if (rc) return rc;
    return launch_fir_axis<T>(h, d_tmp, d_out, Z, Y, X, 2, t[2], fp64, s);
}

// The same march reading the raw uint16 stack: SP:26-36 (float conversion, pedestal, percentile clip) happen on the
// way into the window, so the fp32 path never writes the un-blurred float volume (one 1 GB write + read less at
// 2048 x 2048 x 64)
template <int R>
__global__ void __launch_bounds__(256) prep_fir_z_march_kernel(const uint16_t* __restrict__ in, float* __restrict__ out,
                                                               int Z, size_t plane4, const float* __restrict__ w32,
                                                               int pedestal, const int32_t* __restrict__ status) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= plane4) return;
    const bool clip = st_load(status, ST_HAS_NONZERO) != 0;
    const float p95 = clip ? __int_as_float(st_load(status, ST_P95_BITS)) : 3.0e38f;
    const uint2* src = reinterpret_cast<const uint2*>(in) + p;
    float4* dst = reinterpret_cast<float4*>(out) + p;
    auto load = [&](int z) {
        const uint2 q = __ldg(src + (size_t)clampi(z, 0, Z - 1) * plane4);
        const int v0 = (int)(q.x & 0xffffu) - pedestal, v1 = (int)(q.x >> 16) - pedestal;
        const int v2 = (int)(q.y & 0xffffu) - pedestal, v3 = (int)(q.y >> 16) - pedestal;
        return make_float4(fminf((float)max(v0, 0), p95), fminf((float)max(v1, 0), p95), fminf((float)max(v2, 0), p95),
                           fminf((float)max(v3, 0), p95));
    };
    float w[2 * R + 1];
#pragma unroll
    for (int k = 0; k <= 2 * R; ++k) w[k] = __ldg(w32 + k);
    float4 win[2 * R + 1];
#pragma unroll
    for (int k = 0; k <= 2 * R; ++k) win[k] = load(k - R);
    for (int z0 = 0; z0 < Z; z0 += 2 * R + 1) {
#pragma unroll
        for (int u = 0; u <= 2 * R; ++u) {
            const int z = z0 + u;
            if (z >= Z) break;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k <= 2 * R; ++k) {
                const float4 v = win[(u + k) % (2 * R + 1)];
                a.x = fmaf(w[k], v.x, a.x); a.y = fmaf(w[k], v.y, a.y);
                a.z = fmaf(w[k], v.z, a.z); a.w = fmaf(w[k], v.w, a.w);
            }
            dst[(size_t)z * plane4] = a;
            win[u] = load(z + R + 1);
        }
    }
}

// SP:26-37 in fp32: prepared volume blurred with `sigma`, result in d_out, d_tmp is scratch of the same size.  Takes
// the fused march when it applies, otherwise prepare + the three generic passes (d_tmp holds the prepared volume).
int prepare_and_blur_f32(tsp_handle* h, const uint16_t* d_in, float* d_out, float* d_tmp, int Z, int Y, int X,
                         const double sigma[3], int pedestal, const int32_t* d_status, cudaStream_t s) {
    DeviceTaps t[3];
    for (int a = 0; a < 3; ++a) {
        int rc = get_taps(h, sigma[a], &t[a]);
        if (rc) return rc;
    }
    const size_t plane = (size_t)Y * X;
    const int r = t[0].radius;
    const bool fused = (r == 0 || r == 1 || r == 2 || r == 4) && plane % 4 == 0 &&
                       (reinterpret_cast<uintptr_t>(d_in) & 7) == 0 &&
                       ((reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(d_tmp)) & 15) == 0;
    if (!fused) {
        int rc = launch_prepare(h, d_in, d_tmp, (size_t)Z * plane, pedestal, d_status, s);
        if (rc) return rc;
        return gaussian_blur<float>(h, d_tmp, d_out, d_tmp, Z, Y, X, sigma, false, s);
    }
    const size_t plane4 = plane / 4;
    const unsigned blocks = (unsigned)((plane4 + 255) / 256);
    const bool short_xy = t[1].radius <= 8 && t[2].radius <= 8;
    float* zdst = short_xy ? d_tmp : d_out;          // the fused y/x kernel then goes d_tmp -> d_out
    if (r == 0) prep_fir_z_march_kernel<0><<<blocks, 256, 0, s>>>(d_in, zdst, Z, plane4, t[0].w32, pedestal, d_status);
    else if (r == 1) prep_fir_z_march_kernel<1><<<blocks, 256, 0, s>>>(d_in, zdst, Z, plane4, t[0].w32, pedestal, d_status);
    else if (r == 2) prep_fir_z_march_kernel<2><<<blocks, 256, 0, s>>>(d_in, zdst, Z, plane4, t[0].w32, pedestal, d_status);
    else prep_fir_z_march_kernel<4><<<blocks, 256, 0, s>>>(d_in, zdst, Z, plane4, t[0].w32, pedestal, d_status);
    TSP_LAUNCH_CHECK(h);
    if (short_xy) {
        fir_yx_fused(h, d_tmp, d_out, Z, Y, X, t[1], t[2], s);
        if (cudaGetLastError() != cudaSuccess) {
            set_error("fused y/x pass failed to launch");
            return TSP_ERR_CUDA;
        }
        return TSP_OK;
    }
    int rc = launch_fir_axis<float>(h, d_out, d_tmp, Z, Y, X, 1, t[1], false, s);
    if (rc) return rc;
    return launch_fir_axis<float>(h, d_tmp, d_out, Z, Y, X, 2, t[2], false, s);
}

template int gaussian_blur<float>(tsp_handle*, const float*, float*, float*, int, int, int, const double[3], bool, cudaStream_t);
template int gaussian_blur<uint16_t>(tsp_handle*, const uint16_t*, uint16_t*, uint16_t*, int, int, int, const double[3], bool, cudaStream_t);
template int launch_fir_axis<float>(tsp_handle*, const float*, float*, int, int, int, int, const DeviceTaps&, bool, cudaStream_t);
template int launch_fir_axis<uint16_t>(tsp_handle*, const uint16_t*, uint16_t*, int, int, int, int, const DeviceTaps&, bool, cudaStream_t);

}  // namespace tsp
