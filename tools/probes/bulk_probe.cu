// Microbenchmark: streaming read of a large buffer through per-warp rings of cp.async.bulk copies of ITEM bytes,
// vs plain 16-byte loads.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_probe bulk_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int ITEM, int STAGES, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) ring_kernel(const uint4* __restrict__ src, size_t nitems, unsigned long long* out) {
    extern __shared__ __align__(128) unsigned char ring[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ring_s = smem_u32(ring) + warp * STAGES * ITEM;
    const uint32_t bar_s = smem_u32(ring) + WARPS * STAGES * ITEM + warp * STAGES * 8;
    const uint4* rv = reinterpret_cast<const uint4*>(ring + warp * STAGES * ITEM);
    if (lane == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(bar_s + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    const size_t stride = (size_t)gridDim.x * WARPS;
    size_t issue_it = (size_t)blockIdx.x * WARPS + warp;
    for (int s = 0; s < STAGES - 1; ++s) {
        if (lane == 0 && issue_it < nitems) { mbar_expect_tx(bar_s + 8 * s, ITEM); bulk_load_1d(ring_s + s * ITEM, src + issue_it * (ITEM / 16), ITEM, bar_s + 8 * s); }
        issue_it += stride;
    }
    int stage = 0; uint32_t phase = 0;
    uint32_t acc = 0;
    for (size_t it = (size_t)blockIdx.x * WARPS + warp; it < nitems; it += stride) {
        const int ps = stage == 0 ? STAGES - 1 : stage - 1;
        if (lane == 0 && issue_it < nitems) { mbar_expect_tx(bar_s + 8 * ps, ITEM); bulk_load_1d(ring_s + ps * ITEM, src + issue_it * (ITEM / 16), ITEM, bar_s + 8 * ps); }
        issue_it += stride;
        mbar_wait(bar_s + 8 * stage, phase);
#pragma unroll
        for (int j = 0; j < ITEM / 512; ++j) {
            const uint4 v = rv[stage * (ITEM / 16) + j * 32 + lane];
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
    if (acc == 0x12345678u) atomicAdd(out, 1ull);
}

template <int U>
__global__ void __launch_bounds__(512, 2) ldg_kernel(const uint4* __restrict__ src, size_t nvec, unsigned long long* out) {
    uint32_t acc = 0;
    const size_t chunk = (size_t)512 * U;
    for (size_t c = blockIdx.x; c < nvec / chunk; c += gridDim.x) {
        uint4 v[U];
#pragma unroll
        for (int j = 0; j < U; ++j) v[j] = __ldg(src + c * chunk + j * 512 + threadIdx.x);
#pragma unroll
        for (int j = 0; j < U; ++j) acc += v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
    }
    if (acc == 0x12345678u) atomicAdd(out, 1ull);
}

template <int ITEM, int STAGES, int WARPS>
void run_ring(const uint4* d, size_t bytes, unsigned long long* out, int ctas_per_sm) {
    const int smem = WARPS * STAGES * ITEM + WARPS * STAGES * 8;
    cudaFuncSetAttribute(ring_kernel<ITEM, STAGES, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a);
        ring_kernel<ITEM, STAGES, WARPS><<<148 * ctas_per_sm, WARPS * 32, smem>>>(d, bytes / ITEM, out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    printf("ring item %5d B stages %d warps/CTA %2d CTAs/SM %d smem %6d: %.3f ms  %.0f GB/s  (%s)\n", ITEM, STAGES, WARPS, ctas_per_sm, smem, best,
           bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const size_t bytes = (size_t)1 << 30;
    uint4* d; cudaMalloc(&d, bytes); cudaMemset(d, 1, bytes);
    unsigned long long* out; cudaMalloc(&out, 8); cudaMemset(out, 0, 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int r = 0; r < 2; ++r) {
        float best = 1e9;
        for (int k = 0; k < 5; ++k) {
            cudaEventRecord(a);
            if (r == 0) ldg_kernel<4><<<296, 512>>>(d, bytes / 16, out); else ldg_kernel<8><<<296, 512>>>(d, bytes / 16, out);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        printf("ldg x%d: %.3f ms %.0f GB/s\n", r == 0 ? 4 : 8, best, bytes / best / 1e6);
    }
    run_ring<1024, 4, 16>(d, bytes, out, 2);
    run_ring<2048, 4, 8>(d, bytes, out, 2);
    run_ring<2048, 3, 16>(d, bytes, out, 2);
    run_ring<4096, 3, 8>(d, bytes, out, 2);
    run_ring<4096, 3, 4>(d, bytes, out, 4);
    run_ring<4096, 4, 4>(d, bytes, out, 3);
    run_ring<8192, 3, 4>(d, bytes, out, 2);
    run_ring<8192, 2, 8>(d, bytes, out, 1);
    run_ring<16384, 3, 4>(d, bytes, out, 1);
    run_ring<16384, 2, 4>(d, bytes, out, 1);
    return 0;
}
