// Can a PERSISTENT kernel reach the write rate of a huge grid of short-lived blocks (~7.4 TB/s)?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// (b) persistent CTAs claim 8 KB tiles with an atomic ticket
__global__ void __launch_bounds__(128) k_ticket(char* J, size_t ntiles, unsigned long long* ctr) {
    __shared__ unsigned long long s_t;
    for (;;) {
        if (threadIdx.x == 0) s_t = atomicAdd(ctr, 1ULL);
        __syncthreads();
        const size_t t = s_t;
        __syncthreads();
        if (t >= ntiles) return;
        char* base = J + t * 8192;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            *reinterpret_cast<double2*>(base + ((size_t)u * 128 + threadIdx.x) * 16) = make_double2(0.0, 0.0);
    }
}
// (b2) per-warp ticket, 2 KB tiles per warp visit x4
__global__ void __launch_bounds__(256) k_ticket_warp(char* J, size_t ntiles, unsigned long long* ctr) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(ctr, 1ULL);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= ntiles) return;
        char* base = J + t * 8192 + lane * 16;
#pragma unroll
        for (int u = 0; u < 16; ++u) *reinterpret_cast<double2*>(base + u * 512) = make_double2(0.0, 0.0);
    }
}
// (d) persistent CTAs, TMA bulk stores from a zero buffer in shared memory; one thread issues
template <int KB>
__global__ void __launch_bounds__(128) k_tma(char* J, size_t bytes) {
    extern __shared__ __align__(128) char zbuf[];
    for (int i = threadIdx.x * 16; i < KB * 1024; i += blockDim.x * 16) *reinterpret_cast<double2*>(zbuf + i) = make_double2(0.0, 0.0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t piece = (size_t)KB * 1024;
        for (size_t o = (size_t)blockIdx.x * piece; o + piece <= bytes; o += (size_t)gridDim.x * piece) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(J + o), "r"(smem_u32(zbuf)), "r"((uint32_t)piece) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
// (e) persistent slabs, U independent 16-byte stores per thread per iteration
template <int U>
__global__ void __launch_bounds__(256) k_slabs(char* J, size_t bytes) {
    const size_t slab = ((bytes / gridDim.x) + 4095) & ~(size_t)4095;
    const size_t lo = (size_t)blockIdx.x * slab, hi = lo + slab < bytes ? lo + slab : bytes;
    for (size_t o = lo + (size_t)threadIdx.x * 16; o < hi; o += (size_t)256 * 16 * U) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t q = o + (size_t)u * 256 * 16;
            if (q < hi) *reinterpret_cast<double2*>(J + q) = make_double2(0.0, 0.0);
        }
    }
}
// (f) non-persistent: huge grid, each 256-thread block zeroes `cols` columns of 3656 B like K2b would
__global__ void __launch_bounds__(256) k_colblocks(double* J, int n, int M, int cols_per_block) {
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
}
int main() {
    const size_t ntiles = 368640, bytes = ntiles * 8192;
    char *J, *flush; unsigned long long* ctr;
    cudaMalloc(&J, bytes + 65536); cudaMalloc(&flush, 256u << 20); cudaMalloc(&ctr, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto launch) {
        float sum = 0.f;
        for (int r = 0; r < 13; ++r) {
            cudaMemsetAsync(flush, 0, 256u << 20); cudaMemsetAsync(ctr, 0, 8);
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (r >= 3) sum += ms;
        }
        cudaError_t e = cudaGetLastError();
        printf("%-56s avg %.3f ms %.0f GB/s %s\n", name, sum / 10, bytes / (sum / 10) / 1e6, e ? cudaGetErrorString(e) : "");
    };
    run("(b) CTA ticket, 8 KB tiles, 148*16 CTAs x128", [&] { k_ticket<<<148 * 16, 128>>>(J, ntiles, ctr); });
    run("(b) CTA ticket, 8 KB tiles, 148*4 CTAs x128", [&] { k_ticket<<<148 * 4, 128>>>(J, ntiles, ctr); });
    run("(b2) warp ticket, 8 KB per visit, 444 CTAs x256", [&] { k_ticket_warp<<<444, 256>>>(J, ntiles, ctr); });
    cudaFuncSetAttribute(k_tma<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    cudaFuncSetAttribute(k_tma<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 1024);
    cudaFuncSetAttribute(k_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 1024);
    run("(d) TMA bulk stores 32 KB, 444 CTAs", [&] { k_tma<32><<<444, 128, 32 * 1024>>>(J, bytes); });
    run("(d) TMA bulk stores 8 KB, 444 CTAs", [&] { k_tma<8><<<444, 128, 8 * 1024>>>(J, bytes); });
    run("(d) TMA bulk stores 4 KB, 888 CTAs", [&] { k_tma<4><<<888, 128, 4 * 1024>>>(J, bytes); });
    run("(d) TMA bulk stores 8 KB, 148 CTAs", [&] { k_tma<8><<<148, 128, 8 * 1024>>>(J, bytes); });
    run("(e) slabs U=1, 444 CTAs", [&] { k_slabs<1><<<444, 256>>>(J, bytes); });
    run("(e) slabs U=8, 444 CTAs", [&] { k_slabs<8><<<444, 256>>>(J, bytes); });
    run("(e) slabs U=16, 444 CTAs", [&] { k_slabs<16><<<444, 256>>>(J, bytes); });
    const int n = 201, M = 457, B = 4096;
    run("(f) huge grid, 8 columns per 256-thread block", [&] { k_colblocks<<<(unsigned)((size_t)B * n / 8), 256>>>((double*)J, n, M, 8); });
    run("(f) huge grid, 24 columns per 256-thread block", [&] { k_colblocks<<<(unsigned)((size_t)B * n / 24), 256>>>((double*)J, n, M, 24); });
    cudaDeviceSynchronize();
    return 0;
}
