// Development probe: what does a read-only stream reach on this part?  (4 GiB buffer: no L2 reuse between runs.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/read_probe tools/probes/read_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <int T, int U>
__global__ void __launch_bounds__(T) persistent_read(const uint4* __restrict__ src, size_t nvec, unsigned long long* out) {
    unsigned acc = 0;
    const size_t chunk = (size_t)T * U;
    for (size_t c = blockIdx.x; c < nvec / chunk; c += gridDim.x) {
        uint4 v[U];
#pragma unroll
        for (int j = 0; j < U; ++j) v[j] = __ldg(src + c * chunk + j * T + threadIdx.x);
#pragma unroll
        for (int j = 0; j < U; ++j) acc += v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
    }
    if (acc == 0x12345u) atomicAdd(out, 1ull);
}

// one CTA per contiguous tile (non-persistent, like an elementwise library kernel)
template <int T, int U>
__global__ void __launch_bounds__(T) tiled_read(const uint4* __restrict__ src, size_t nvec, unsigned long long* out) {
    unsigned acc = 0;
    const size_t base = (size_t)blockIdx.x * T * U;
    uint4 v[U];
#pragma unroll
    for (int j = 0; j < U; ++j) v[j] = __ldg(src + base + j * T + threadIdx.x);
#pragma unroll
    for (int j = 0; j < U; ++j) acc += v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
    if (acc == 0x12345u) atomicAdd(out, 1ull);
}

template <int T, int U>
__global__ void __launch_bounds__(T) tiled_copy(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t nvec) {
    const size_t base = (size_t)blockIdx.x * T * U;
    uint4 v[U];
#pragma unroll
    for (int j = 0; j < U; ++j) v[j] = __ldg(src + base + j * T + threadIdx.x);
#pragma unroll
    for (int j = 0; j < U; ++j) dst[base + j * T + threadIdx.x] = v[j];
}

template <typename F>
float best_ms(F launch, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e9f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a);
        launch(r);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const size_t total = (size_t)4 << 30, win = (size_t)512 << 20;
    uint4* d;
    if (cudaMalloc(&d, total) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(d, 1, total);
    unsigned long long* out;
    cudaMalloc(&out, 8);
    cudaMemset(out, 0, 8);
    const size_t wvec = win / 16;
    auto at = [&](int r) { return d + (size_t)(r % 8) * wvec; };       // a fresh 512 MiB window every run
    float ms;
#define REPORT(name, bytes) printf("%-44s %.3f ms  %.0f GB/s\n", name, ms, (double)(bytes) / ms / 1e6)
    ms = best_ms([&](int r) { persistent_read<512, 4><<<296, 512>>>(at(r), wvec, out); }, 8); REPORT("persistent 296x512 U4, 512 MiB", win);
    ms = best_ms([&](int r) { persistent_read<512, 8><<<296, 512>>>(at(r), wvec, out); }, 8); REPORT("persistent 296x512 U8, 512 MiB", win);
    ms = best_ms([&](int r) { persistent_read<256, 4><<<148 * 8, 256>>>(at(r), wvec, out); }, 8); REPORT("persistent 1184x256 U4, 512 MiB", win);
    ms = best_ms([&](int r) { persistent_read<1024, 4><<<296, 1024>>>(at(r), wvec, out); }, 8); REPORT("persistent 296x1024 U4, 512 MiB", win);
    ms = best_ms([&](int r) { tiled_read<256, 4><<<(unsigned)(wvec / 1024), 256>>>(at(r), wvec, out); }, 8); REPORT("tiled 256 U4 (16 KB / CTA), 512 MiB", win);
    ms = best_ms([&](int r) { tiled_read<512, 4><<<(unsigned)(wvec / 2048), 512>>>(at(r), wvec, out); }, 8); REPORT("tiled 512 U4 (32 KB / CTA), 512 MiB", win);
    ms = best_ms([&](int r) { tiled_read<256, 8><<<(unsigned)(wvec / 2048), 256>>>(at(r), wvec, out); }, 8); REPORT("tiled 256 U8 (32 KB / CTA), 512 MiB", win);
    ms = best_ms([&](int r) { persistent_read<512, 4><<<296, 512>>>(d, total / 16, out); }, 3); REPORT("persistent 296x512 U4, 4 GiB", total);
    ms = best_ms([&](int r) { tiled_read<256, 4><<<(unsigned)(total / 16 / 1024), 256>>>(d, total / 16, out); }, 3); REPORT("tiled 256 U4, 4 GiB", total);
    ms = best_ms([&](int r) { tiled_copy<256, 4><<<(unsigned)(wvec / 1024), 256>>>(at(r), d + (size_t)((r + 4) % 8) * wvec, wvec); }, 8); REPORT("tiled copy 256 U4, 512 MiB (r+w bytes)", 2 * win);
    ms = best_ms([&](int r) { cudaMemcpyAsync(d + (size_t)((r + 4) % 8) * wvec, at(r), win, cudaMemcpyDeviceToDevice); }, 8); REPORT("cudaMemcpy D2D 512 MiB (r+w bytes)", 2 * win);
    ms = best_ms([&](int r) { cudaMemsetAsync(at(r), 0, win); }, 8); REPORT("cudaMemset 512 MiB (w bytes)", win);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
