// Does the column-pattern fill depend on resident warps per SM?  Huge grid, 8 columns of 3656 B per
// 256-thread block (tools/fill_persist.cu variant f), with dynamic shared memory used only to cap
// the number of resident blocks per SM.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__global__ void __launch_bounds__(256) k_colblocks(double* J, int M, int cols_per_block) {
    extern __shared__ double pad[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t c0 = (size_t)blockIdx.x * cols_per_block;
    for (int c = warp; c < cols_per_block; c += 8) {
        double* dst = J + (c0 + c) * M;
        const unsigned hj = (unsigned)((reinterpret_cast<uintptr_t>(dst) >> 3) & 1);
        const unsigned n2 = ((unsigned)M - hj) >> 1;
        double2* p2 = reinterpret_cast<double2*>(dst + hj);
        for (unsigned i = lane; i < n2; i += 32) p2[i] = make_double2(0.0, 0.0);
        if (lane == 0 && hj) dst[0] = 0.0;
        if (lane == 1 && ((M - hj) & 1)) dst[M - 1] = 0.0;
    }
    if (cols_per_block < 0) pad[threadIdx.x] = 0.0;
}
int main() {
    const int n = 201, M = 457, B = 4096;
    const size_t total = (size_t)B * n * M;
    double* J; char* flush;
    cudaMalloc(&J, total * 8 + 256); cudaMalloc(&flush, 256u << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(k_colblocks, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int cols : {8, 24, 201}) for (int smem_kb : {0, 28, 44, 75, 110, 200}) {
        float sum = 0.f;
        for (int r = 0; r < 13; ++r) {
            cudaMemsetAsync(flush, 0, 256u << 20);
            cudaEventRecord(e0);
            k_colblocks<<<(unsigned)((size_t)B * n / cols), 256, smem_kb * 1024>>>(J, M, cols);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (r >= 3) sum += ms;
        }
        int blocks = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_colblocks, 256, smem_kb * 1024);
        printf("%3d columns per block, %3d KB smem -> %d blocks (%2d warps) per SM: %.3f ms %.0f GB/s\n", cols, smem_kb, blocks,
               blocks * 8, sum / 10, total * 8 / (sum / 10) / 1e6);
    }
    return 0;
}
