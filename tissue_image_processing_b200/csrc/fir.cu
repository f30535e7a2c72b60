// Direct separable Gaussian FIR passes = scipy.ndimage.gaussian_filter(mode='nearest'), the only
// arithmetic primitive of the reference (basic_image_manipulations.py:373-390).  One kernel per
// axis; each pass stores its output dtype before the next pass like scipy does (float32 rounds,
// uint16 truncates).  Two accumulation policies:
//   fp32  - sequential FMA accumulation, register blocked (TSP_MODE_EXACT)
//   fp64  - scipy's own summation order: centre tap first, then the symmetric pairs from the
//           outermost inwards, (a + b) * w added with separate mul/add (TSP_MODE_BITEXACT)
#include <math.h>

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

template <typename T>
int launch_fir_axis(tsp_handle* h, const T* d_in, T* d_out, int Z, int Y, int X, int axis,
                    const DeviceTaps& taps, bool fp64, cudaStream_t s) {
    const int r = taps.radius;
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
    int rc = launch_fir_axis<T>(h, d_in, d_out, Z, Y, X, 0, t[0], fp64, s);
    if (rc) return rc;
    rc = launch_fir_axis<T>(h, d_out, d_tmp, Z, Y, X, 1, t[1], fp64, s);
    if (rc) return rc;
    return launch_fir_axis<T>(h, d_tmp, d_out, Z, Y, X, 2, t[2], fp64, s);
}

template int gaussian_blur<float>(tsp_handle*, const float*, float*, float*, int, int, int, const double[3], bool, cudaStream_t);
template int gaussian_blur<uint16_t>(tsp_handle*, const uint16_t*, uint16_t*, uint16_t*, int, int, int, const double[3], bool, cudaStream_t);
template int launch_fir_axis<float>(tsp_handle*, const float*, float*, int, int, int, int, const DeviceTaps&, bool, cudaStream_t);
template int launch_fir_axis<uint16_t>(tsp_handle*, const uint16_t*, uint16_t*, int, int, int, int, const DeviceTaps&, bool, cudaStream_t);

}  // namespace tsp
