// surface_projection_m (reference surface_proj_m.py:14-47, 81-100): uint16 Gaussian blur (5,5,3) with
// per-pass truncation -> block mean / variance over (1,bin,bin) blocks (zero padded to a multiple,
// like skimage.measure.block_reduce) -> nearest-neighbour upsample -> argmax over z -> pick the
// blurred voxel (np.choose).
#include "common.cuh"

namespace tsp {

// one thread per (z, by, bx) block; float64 statistics like numpy (sums of uint16 are exact)
__global__ void block_stat_kernel(const uint16_t* __restrict__ vol, double* __restrict__ stat, int Z, int Y,
                                  int X, int bin, int by_n, int bx_n, int method) {
    const size_t n = (size_t)Z * by_n * bx_n;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const double cnt = (double)bin * (double)bin;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int bx = (int)(i % bx_n), by = (int)((i / bx_n) % by_n), z = (int)(i / ((size_t)bx_n * by_n));
        const uint16_t* plane = vol + (size_t)z * Y * X;
        unsigned long long sum = 0;
        for (int dy = 0; dy < bin; ++dy) {
            const int y = by * bin + dy;
            if (y >= Y) break;
            for (int dx = 0; dx < bin; ++dx) {
                const int x = bx * bin + dx;
                if (x >= X) break;
                sum += plane[(size_t)y * X + x];
            }
        }
        const double mean = (double)sum / cnt;
        if (method == 0) {
            stat[i] = mean;
        } else {
            double acc = 0.0;
            for (int dy = 0; dy < bin; ++dy) {
                const int y = by * bin + dy;
                for (int dx = 0; dx < bin; ++dx) {
                    const int x = bx * bin + dx;
                    const double v = (y < Y && x < X) ? (double)plane[(size_t)y * X + x] : 0.0;   // zero padding
                    const double d = v - mean;
                    acc += d * d;
                }
            }
            stat[i] = acc / cnt;
        }
    }
}

__global__ void choose_kernel(const uint16_t* __restrict__ vol, const double* __restrict__ stat,
                              uint16_t* __restrict__ out, int Z, int Y, int X, int bin, int by_n, int bx_n) {
    const size_t plane = (size_t)Y * X;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < plane; p += stride) {
        const int x = (int)(p % X), y = (int)(p / X);
        const size_t cell = (size_t)(y / bin) * bx_n + (x / bin);
        double best = stat[cell];
        int bz = 0;
        for (int z = 1; z < Z; ++z) {
            const double v = stat[(size_t)z * by_n * bx_n + cell];
            if (v > best) {
                best = v;
                bz = z;
            }
        }
        out[p] = vol[(size_t)bz * plane + p];
    }
}

int launch_project_m(tsp_handle* h, const uint16_t* d_channel, uint16_t* d_out, int Z, int Y, int X,
                     int method, int bin, void* d_ws, cudaStream_t s) {
    const size_t vox = (size_t)Z * Y * X;
    uint16_t* blurred = (uint16_t*)d_ws;
    uint16_t* tmp = blurred + align_up(vox, 128);
    const int by_n = (Y + bin - 1) / bin, bx_n = (X + bin - 1) / bin;
    double* stat = (double*)(tmp + align_up(vox, 128));
    const double sig[3] = {5.0, 5.0, 3.0};                         // SPM:18
    int rc = gaussian_blur<uint16_t>(h, d_channel, blurred, tmp, Z, Y, X, sig, true, s);
    if (rc) return rc;
    const size_t ncell = (size_t)Z * by_n * bx_n;
    size_t blocks = (ncell + 127) / 128;
    if (blocks > (size_t)h->sm_count * 32) blocks = (size_t)h->sm_count * 32;
    block_stat_kernel<<<(int)blocks, 128, 0, s>>>(blurred, stat, Z, Y, X, bin, by_n, bx_n, method);
    TSP_LAUNCH_CHECK(h);
    blocks = ((size_t)Y * X + 255) / 256;
    if (blocks > (size_t)h->sm_count * 32) blocks = (size_t)h->sm_count * 32;
    choose_kernel<<<(int)blocks, 256, 0, s>>>(blurred, stat, d_out, Z, Y, X, bin, by_n, bx_n);
    TSP_LAUNCH_CHECK(h);
    return TSP_OK;
}

}  // namespace tsp
