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
    return 0;
}
