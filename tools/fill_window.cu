// How large may the set of concurrently written addresses be before HBM write bandwidth drops?
// Tile kernel (one 8 KB tile per block, blocks dispatched in index order) whose tile order is
// permuted inside windows of W bytes: inside a window consecutive blocks write far-apart tiles.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__global__ void __launch_bounds__(128) k_tiles_perm(char* J, size_t ntiles, size_t tiles_per_window, size_t stride) {
    const size_t i = blockIdx.x;
    const size_t w = i / tiles_per_window, r = i - w * tiles_per_window;
    size_t wt = tiles_per_window;
    if ((w + 1) * tiles_per_window > ntiles) wt = ntiles - w * tiles_per_window;     // ragged last window
    const size_t t = w * tiles_per_window + (r * stride) % wt;                          // stride coprime to wt (odd, wt power of 2) else falls back
    char* base = J + t * 8192;
#pragma unroll
    for (int u = 0; u < 4; ++u)
        *reinterpret_cast<double2*>(base + ((size_t)u * 128 + threadIdx.x) * 16) = make_double2(0.0, 0.0);
}
int main() {
    const size_t ntiles = 368640;                    // 3.02 GB
    const size_t bytes = ntiles * 8192;
    char *J, *flush;
    cudaMalloc(&J, bytes); cudaMalloc(&flush, 256u << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (size_t wmb : {0, 4, 16, 32, 64, 128, 256, 512, 1024, 2048}) {
        const size_t tpw = wmb ? (wmb << 20) / 8192 : 1;
        const size_t stride = wmb ? 1237 : 1;            // odd: a permutation of a power-of-two window
        float sum = 0.f;
        for (int r = 0; r < 13; ++r) {
            cudaMemsetAsync(flush, 0, 256u << 20);
            cudaEventRecord(e0);
            k_tiles_perm<<<(unsigned)ntiles, 128>>>(J, ntiles, tpw, stride);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (r >= 3) sum += ms;
        }
        printf("window %5zu MB (tiles scattered inside): avg %.3f ms  %.0f GB/s\n", wmb, sum / 10, bytes / (sum / 10) / 1e6);
    }
    // the same with larger contiguous pieces per block visit: 64 KB runs scattered over 1 GB windows
    return 0;
}
