// Write-pattern microbenchmark behind the K2 redesign: how fast can 444 persistent CTAs zero a
// [B, n, M] fp64 Jacobian (Goddard-50: n = 201, M = 457) with different store layouts?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fill_patterns.cu -o /tmp/fill_patterns && /tmp/fill_patterns
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ int laneid() { int l; asm volatile("mov.u32 %0, %%laneid;" : "=r"(l)); return l; }

// 0: plain linear fill, grid-stride 16-byte stores
__global__ void __launch_bounds__(256) k_linear(double* J, size_t total) {
    double2* p = reinterpret_cast<double2*>(J);
    const size_t n2 = total / 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_double2(0.0, 0.0);
}

// zero [dst, dst + count) doubles with one warp: 8-byte head/tail, 16-byte body (K2's column stream)
__device__ __forceinline__ void warp_zero(double* dst, unsigned count, int lane) {
    const unsigned hj = (unsigned)((reinterpret_cast<uintptr_t>(dst) >> 3) & 1);
    const unsigned nbytes = ((count - hj) & ~1u) * 8u;
    char* g = reinterpret_cast<char*>(dst + hj) + lane * 16;
    const double2 z2 = make_double2(0.0, 0.0);
    unsigned left = nbytes;
    for (; left >= 2048u; left -= 2048u, g += 2048) {
        *reinterpret_cast<double2*>(g) = z2;
        *reinterpret_cast<double2*>(g + 512) = z2;
        *reinterpret_cast<double2*>(g + 1024) = z2;
        *reinterpret_cast<double2*>(g + 1536) = z2;
    }
    const unsigned mine = lane * 16u;
    if (mine < left) *reinterpret_cast<double2*>(g) = z2;
    if (mine + 512u < left) *reinterpret_cast<double2*>(g + 512) = z2;
    if (mine + 1024u < left) *reinterpret_cast<double2*>(g + 1024) = z2;
    if (mine + 1536u < left) *reinterpret_cast<double2*>(g + 1536) = z2;
    if (lane == 0 && hj) dst[0] = 0.0;
    if (lane == 1 && ((count - hj) & 1)) dst[count - 1] = 0.0;
}

// 1: K2 today -- persistent CTAs, item = instance, one warp per column, round robin
__global__ void __launch_bounds__(256, 3) k_columns(double* J, int B, int n, int M) {
    const int lane = laneid(), warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        double* Jb = J + (size_t)b * n * M;
        for (int c = warp; c < n; c += nw) { warp_zero(Jb + (size_t)c * M, M, lane); __syncwarp(); }
    }
}

// 2: each warp owns a contiguous range of columns and zeroes it `K` columns at a time as one span
__global__ void __launch_bounds__(256, 3) k_warp_ranges(double* J, int B, int n, int M, int K) {
    const int lane = laneid(), warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        double* Jb = J + (size_t)b * n * M;
        const int c0 = (int)((long)n * warp / nw), c1 = (int)((long)n * (warp + 1) / nw);
        for (int c = c0; c < c1; c += K) {
            const int k = min(K, c1 - c);
            warp_zero(Jb + (size_t)c * M, (unsigned)(k * M), lane);
            __syncwarp();
        }
    }
}

// 3: the CTA zeroes chunks of K columns together (aligned linear stores), one barrier per chunk
__global__ void __launch_bounds__(256, 3) k_cta_chunks(double* J, int B, int n, int M, int K) {
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        double* Jb = J + (size_t)b * n * M;
        for (int c = 0; c < n; c += K) {
            const int k = min(K, n - c);
            double* dst = Jb + (size_t)c * M;
            const unsigned count = (unsigned)(k * M);
            const unsigned hj = (unsigned)((reinterpret_cast<uintptr_t>(dst) >> 3) & 1);
            const unsigned n2 = (count - hj) >> 1;
            double2* p2 = reinterpret_cast<double2*>(dst + hj);
            for (unsigned i = threadIdx.x; i < n2; i += blockDim.x) p2[i] = make_double2(0.0, 0.0);
            if (threadIdx.x == 0 && hj) dst[0] = 0.0;
            if (threadIdx.x == 1 && ((count - hj) & 1)) dst[count - 1] = 0.0;
            __syncthreads();
        }
    }
}

// 4: like 2, but the spans are cut at 128-byte aligned addresses instead of column boundaries
__global__ void __launch_bounds__(256, 3) k_warp_ranges_aligned(double* J, int B, int n, int M, int K) {
    const int lane = laneid(), warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        double* Jb = J + (size_t)b * n * M;
        const int c0 = (int)((long)n * warp / nw), c1 = (int)((long)n * (warp + 1) / nw);
        double* cur = Jb + (size_t)c0 * M;
        double* const end = Jb + (size_t)c1 * M;
        for (int c = c0; c < c1; c += K) {
            const int k = min(K, c1 - c);
            double* stop = Jb + (size_t)(c + k) * M;                      // must be covered
            if (c + k < c1) stop = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(stop) + 127) & ~(uintptr_t)127);
            if (stop > end) stop = end;
            warp_zero(cur, (unsigned)(stop - cur), lane);
            cur = stop;
            __syncwarp();
        }
    }
}

int main() {
    const int B = 4096, n = 201, M = 457;
    const size_t total = (size_t)B * n * M;
    double* J;
    CK(cudaMalloc(&J, total * 8 + 256));
    char* flush;
    CK(cudaMalloc(&flush, 256u << 20));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto launch) {
        float best = 1e9f, sum = 0.f;
        for (int r = 0; r < 13; ++r) {
            cudaMemsetAsync(flush, 0, 256u << 20);
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (r >= 3) { best = ms < best ? ms : best; sum += ms; }
        }
        printf("%-44s avg %.3f ms  min %.3f ms  %.0f GB/s\n", name, sum / 10, best, total * 8 / (sum / 10) / 1e6);
        return 0;
    };
    const int grid = 444;
    run("0 linear fill (148*8 CTAs)", [&] { k_linear<<<148 * 8, 256>>>(J, total); });
    run("0 linear fill (444 CTAs)", [&] { k_linear<<<444, 256>>>(J, total); });
    run("1 warp per column, M=457 (K2 today)", [&] { k_columns<<<grid, 256>>>(J, B, n, M); });
    run("1 warp per column, M=456 (16B-aligned cols)", [&] { k_columns<<<grid, 256>>>(J, B, n, 456); });
    run("1 warp per column, M=448 (128B-aligned cols)", [&] { k_columns<<<grid, 256>>>(J, B, n, 448); });
    for (int K : {1, 2, 4, 8}) {
        char nm[96]; snprintf(nm, sizeof nm, "2 warp ranges, spans of %d columns", K);
        run(nm, [&] { k_warp_ranges<<<grid, 256>>>(J, B, n, M, K); });
    }
    for (int K : {1, 2, 4, 8}) {
        char nm[96]; snprintf(nm, sizeof nm, "4 warp ranges, 128B-aligned cuts, %d columns", K);
        run(nm, [&] { k_warp_ranges_aligned<<<grid, 256>>>(J, B, n, M, K); });
    }
    for (int K : {8, 16, 32, 201}) {
        char nm[96]; snprintf(nm, sizeof nm, "3 CTA chunks of %d columns + barrier", K);
        run(nm, [&] { k_cta_chunks<<<grid, 256>>>(J, B, n, M, K); });
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
