// Development probe: issue rate of scalar FFMA vs packed FFMA2 on this part (sm_100a).
//   nvcc -arch=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cuda_runtime.h>
#include <stdio.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float a, float b) {
    float2 acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = make_float2(threadIdx.x * 0.001f + j, j * 0.5f);
    float2 w = make_float2(a, a), v = make_float2(b, b + 1e-3f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (MODE == 0) {                       // 32 scalar FFMA (3 register operands)
                acc[j].x = fmaf(v.x, w.x, acc[j].x);
                acc[j].y = fmaf(v.y, w.y, acc[j].y);
            } else {                               // 16 packed FFMA2
                acc[j] = __ffma2_rn(v, w, acc[j]);
            }
        }
        v.x += 1e-7f;                              // keep the loop from being collapsed
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += acc[j].x + acc[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// MODE 2: the register-blocked line filter of fir.cu (line_fir16): 16 float2 accumulators, a 16-entry circular window
// of (w, w) pairs and one input pair per step, both from shared memory - 2 LDS.64 + 16 FFMA2 per input pair.
// `ctas_per_sm` x 256 threads per SM: how close this very instruction pattern gets to the FFMA2 rate above.
__global__ void __launch_bounds__(256) probe_line(float* out, int iters, int nblk) {
    __shared__ float2 tile[(16 * 9) * 32];
    __shared__ float2 wsh[16 * 17];
    for (int i = threadIdx.x; i < 16 * 9 * 32; i += 256) tile[i] = make_float2(1e-3f * (i & 63), 2e-3f * (i & 31));
    for (int i = threadIdx.x; i < 16 * 17; i += 256) wsh[i] = make_float2(1.f / (1 + i), 1.f / (1 + i));
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float2 acc[16], wc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        acc[j] = make_float2(0.f, 0.f);
        wc[j] = make_float2(0.f, 0.f);
    }
    for (int it = 0; it < iters; ++it) {
        const float2* col = tile + lane;
        for (int ii = 0; ii < 16 * nblk; ii += 16) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                wc[u] = wsh[ii + u];
                const float2 v = col[((ii + u) & 127) * 32];
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] = __ffma2_rn(v, wc[(u - j + 16) % 16], acc[j]);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += acc[j].x + acc[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int blocks = prop.multiProcessorCount * 8, iters = 20000;
    float* out;
    cudaMalloc(&out, blocks * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) probe<0><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
            else probe<1><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double fma = (double)blocks * 256 * iters * 32;
            if (rep) printf("%s: %.3f ms, %.1f TFMA/s (%.1f TFLOP/s), %d SMs\n", mode ? "FFMA2" : "FFMA ", ms, fma / ms / 1e9,
                            2 * fma / ms / 1e9, prop.multiProcessorCount);
        }
    }
    for (int ctas = 1; ctas <= 4; ++ctas) {
        const int nb = prop.multiProcessorCount * ctas, it2 = 400, nblk = 16;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            probe_line<<<nb, 256>>>(out, it2, nblk);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double fma = (double)nb * 256 * it2 * nblk * 16 * 32;
            if (rep) printf("line filter pattern, %d x 256 threads per SM: %.3f ms, %.1f TFMA/s\n", ctas, ms, fma / ms / 1e9);
        }
    }
    return 0;
}
