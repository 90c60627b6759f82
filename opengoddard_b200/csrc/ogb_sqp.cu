// ogb_sqp.cu -- the batched SLSQP step on the GPU ("device-side batched QP").
//
// One thread block advances one problem instance through one reverse-communication step of Kraft's SLSQP
// (csrc/ogb_sqp.h: BFGS update of L D L', the QP as LSQ -> LSEI -> LSI -> LDP -> NNLS, the L1 merit line search and
// the convergence tests), reading the constraint values and the PACKED Jacobian the sweep kernel just wrote
// (ogb_eval_sparse / ogb_eval_exact).  Blocks take instances from an atomic ticket, so instances in a cheap
// phase (line search) do not hold up the ones solving a QP.  Persistent per-instance state (x0, s, multipliers,
// L D L') and the per-block QP matrices live in HBM / L2; nothing comes back to the host but one int per instance.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "ogb200.h"
#include "ogb_sqp_host.h"

extern "C" int ogb_set_error_message(const char* msg);      // ogb_kernels.cu (thread-local last error)

#define OGS_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            char buf_[256];                                                                  \
            snprintf(buf_, sizeof buf_, "%s: %s", #call, cudaGetErrorString(e_));            \
            return ogb_set_error_message(buf_);                                              \
        }                                                                                    \
    } while (0)

#define OGS_THREADS 256

// a thread block as the cooperating group of ogb_sqp.h
struct OgsBlock {
    int tid, nthr, warp, nwarps, lane, wsize;
    double* red;
    int* redi;
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ double wsum(double v) {
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ __forceinline__ double sum(double v) {
        v = wsum(v);
        if (lane == 0) red[warp] = v;
        __syncthreads();
        double t = 0.0;
        for (int w = 0; w < nwarps; ++w) t += red[w];
        __syncthreads();
        return t;
    }
    __device__ __forceinline__ double max(double v) {
        for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) red[warp] = v;
        __syncthreads();
        double t = red[0];
        for (int w = 1; w < nwarps; ++w) t = fmax(t, red[w]);
        __syncthreads();
        return t;
    }
    // the largest v over the block; among equal v the smallest (first) or largest index; i < 0 = no candidate
    __device__ __forceinline__ static bool better(double v, int i, double v2, int i2, bool first) {
        if (i2 < 0) return false;
        if (i < 0) return true;
        if (v2 != v) return v2 > v;
        return first ? i2 < i : i2 > i;
    }
    __device__ __forceinline__ void argbest(double& v, int& i, bool first) {
        for (int o = 16; o; o >>= 1) {
            const double v2 = __shfl_xor_sync(0xffffffffu, v, o);
            const int i2 = __shfl_xor_sync(0xffffffffu, i, o);
            if (better(v, i, v2, i2, first)) { v = v2; i = i2; }
        }
        if (lane == 0) { red[warp] = v; redi[warp] = i; }
        __syncthreads();
        double bv = red[0];
        int bi = redi[0];
        for (int w = 1; w < nwarps; ++w)
            if (better(bv, bi, red[w], redi[w], first)) { bv = red[w]; bi = redi[w]; }
        __syncthreads();
        v = bv;
        i = bi;
    }
    __device__ __forceinline__ long long clock() { return clock64(); }
    static constexpr bool kLanes32 = true;
    double* wb;                         // dynamic shared memory: `wrows` row buffers per warp
    int wstride, wrows, flags;
    __device__ __forceinline__ double* wbuf() { return wb + (size_t)warp * wrows * wstride; }
    __device__ __forceinline__ void wsync() { __syncwarp(); }
    __device__ __forceinline__ double wmax(double v) {
        for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
};

// MINB: resident blocks per SM the register allocation aims at (2: no spills; 3, 4: more warps to hide latency)
template <int MINB>
__global__ void __launch_bounds__(OGS_THREADS, MINB)
ogb_sqp_step_kernel(OgsShape S, double* X, const double* C, const double* vals, double* state, double* scratch, int B,
                    int* ticket, int* mode_out, int wrows, int flags) {
    extern __shared__ double s_rows[];
    __shared__ double red[32];
    __shared__ int redi[32];
    __shared__ int s_item;
    OgsBlock cx;
    cx.wb = s_rows;
    cx.wstride = (S.n1 + 3) & ~3;
    cx.wrows = wrows;
    cx.flags = flags;
    cx.tid = threadIdx.x; cx.nthr = blockDim.x; cx.warp = threadIdx.x >> 5; cx.nwarps = blockDim.x >> 5;
    cx.lane = threadIdx.x & 31; cx.wsize = 32; cx.red = red; cx.redi = redi;
    double* W = scratch + (size_t)blockIdx.x * S.scratch_doubles;
    while (true) {
        if (threadIdx.x == 0) s_item = atomicAdd(ticket, 1);
        __syncthreads();
        const int b = s_item;
        __syncthreads();
        if (b >= B) break;
        OgsInst I{&S, X + (size_t)b * S.n, C + (size_t)b * S.M, vals + (size_t)b * S.nnz,
                  state + (size_t)b * S.state_doubles, W};
        ogs_step(cx, I);
        if (threadIdx.x == 0) mode_out[b] = (int)I.st[S.o_sc + OGS_MODE];
    }
}

__global__ void ogb_sqp_reset_kernel(OgsShape S, double* state, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double* sc = state + (size_t)b * S.state_doubles + S.o_sc;
    for (int k = 0; k < OGS_NSC; ++k) sc[k] = 0.0;
    sc[OGS_MODE] = OGS_MODE_START;
}

struct OgbDeviceSqp {
    OgsHostTables T;
    int device = 0, sm_count = 0, max_batch = 0, blocks = 0, threads = OGS_THREADS;
    int *colptr_d = nullptr, *prow_d = nullptr, *rowptr_d = nullptr, *rcol_d = nullptr, *rpos_d = nullptr,
        *blo_d = nullptr, *bhi_d = nullptr, *ticket_d = nullptr, *mode_d = nullptr;
    double *xl_d = nullptr, *xu_d = nullptr, *state_d = nullptr, *scratch_d = nullptr;
    long long launches = 0;
    size_t smem = 0;
    int wrows = 1, minb = 2, flags = 1;
    const void* fn = nullptr;
};

template <class T>
static cudaError_t upload(T** dst, const std::vector<T>& src) {
    cudaError_t e = cudaMalloc((void**)dst, std::max<size_t>(1, src.size()) * sizeof(T));
    if (e != cudaSuccess) return e;
    if (src.empty()) return cudaSuccess;
    return cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
}

extern "C" {

void ogb_sqp_destroy(void* h) {
    OgbDeviceSqp* q = (OgbDeviceSqp*)h;
    if (!q) return;
    cudaFree(q->colptr_d); cudaFree(q->prow_d); cudaFree(q->rowptr_d); cudaFree(q->rcol_d); cudaFree(q->rpos_d);
    cudaFree(q->blo_d); cudaFree(q->bhi_d); cudaFree(q->ticket_d); cudaFree(q->mode_d); cudaFree(q->xl_d);
    cudaFree(q->xu_d); cudaFree(q->state_d); cudaFree(q->scratch_d);
    delete q;
}

void* ogb_sqp_create(int nvars, int m, int meq, int nnz, const int32_t* colptr_h, const int32_t* prow_h, const double* xl_h,
                     const double* xu_h, double acc, int maxiter, int max_batch) {
    if (max_batch <= 0) { ogb_set_error_message("ogb_sqp_create: max_batch must be positive"); return nullptr; }
    OgbDeviceSqp* q = new OgbDeviceSqp();
    std::string err;
    if (!ogs_build_tables(nvars, m, meq, nnz, colptr_h, prow_h, xl_h, xu_h, acc, maxiter, q->T, &err)) {
        ogb_set_error_message(err.c_str());
        delete q;
        return nullptr;
    }
    OgsShape& S = q->T.S;
    if (S.mineq + S.nlo + S.nhi <= 0) {
        ogb_set_error_message("ogb_sqp_create: a problem without inequality constraints and bounds is not supported");
        delete q;
        return nullptr;
    }
    cudaDeviceProp prop;
    bool ok = cudaGetDevice(&q->device) == cudaSuccess && cudaGetDeviceProperties(&prop, q->device) == cudaSuccess;
    if (ok) {
        q->sm_count = prop.multiProcessorCount;
        q->max_batch = max_batch;
        int per_sm = 0;
        if (const char* ev = getenv("OGB200_SQP_THREADS")) {          // block width of the step kernel (experiments)
            const int v = atoi(ev);
            if (v >= 32 && v <= OGS_THREADS && v % 32 == 0) q->threads = v;
        }
        // rows a warp takes through the reflections at once (4, 2 or 1): as many as fit in ~96 KB per block
        const size_t per_row = (size_t)(q->threads / 32) * ((S.n1 + 3) & ~3) * sizeof(double);
        q->wrows = per_row * 4 <= 96 * 1024 ? 4 : (per_row * 2 <= 96 * 1024 ? 2 : 1);
        if (const char* ev = getenv("OGB200_SQP_WROWS")) {
            const int v = atoi(ev);
            if (v == 1 || v == 2 || v == 4) q->wrows = v;
        }
        q->smem = per_row * q->wrows;
        if (const char* ev = getenv("OGB200_SQP_FLAGS")) q->flags = atoi(ev);   // bit 0: register-resident reflectors
        if (const char* ev = getenv("OGB200_SQP_MINBLOCKS")) {
            const int v = atoi(ev);
            if (v >= 2 && v <= 4) q->minb = v;
        }
        q->fn = q->minb == 4 ? (const void*)ogb_sqp_step_kernel<4>
                             : (q->minb == 3 ? (const void*)ogb_sqp_step_kernel<3> : (const void*)ogb_sqp_step_kernel<2>);
        ok = q->smem <= 200 * 1024 &&
             cudaFuncSetAttribute(q->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q->smem) == cudaSuccess &&
             cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, q->fn, q->threads, q->smem) == cudaSuccess;
        q->blocks = std::max(1, std::min(max_batch, q->sm_count * std::max(1, per_sm)));
    }
    ok = ok && upload(&q->colptr_d, q->T.colptr) == cudaSuccess && upload(&q->prow_d, q->T.prow) == cudaSuccess &&
         upload(&q->rowptr_d, q->T.rowptr) == cudaSuccess && upload(&q->rcol_d, q->T.rcol) == cudaSuccess &&
         upload(&q->rpos_d, q->T.rpos) == cudaSuccess && upload(&q->blo_d, q->T.blo) == cudaSuccess &&
         upload(&q->bhi_d, q->T.bhi) == cudaSuccess && upload(&q->xl_d, q->T.xl) == cudaSuccess &&
         upload(&q->xu_d, q->T.xu) == cudaSuccess &&
         cudaMalloc((void**)&q->ticket_d, sizeof(int)) == cudaSuccess &&
         cudaMalloc((void**)&q->mode_d, (size_t)max_batch * sizeof(int)) == cudaSuccess &&
         cudaMalloc((void**)&q->state_d, (size_t)max_batch * S.state_doubles * sizeof(double)) == cudaSuccess &&
         cudaMalloc((void**)&q->scratch_d, (size_t)q->blocks * S.scratch_doubles * sizeof(double)) == cudaSuccess;
    if (!ok) {
        char buf[200];
        snprintf(buf, sizeof buf, "ogb_sqp_create: %s", cudaGetErrorString(cudaGetLastError()));
        ogb_set_error_message(buf);
        ogb_sqp_destroy(q);
        return nullptr;
    }
    S.colptr = q->colptr_d; S.prow = q->prow_d; S.rowptr = q->rowptr_d; S.rcol = q->rcol_d; S.rpos = q->rpos_d;
    S.blo = q->blo_d; S.bhi = q->bhi_d; S.xl = q->xl_d; S.xu = q->xu_d;
    return q;
}

// bytes of device memory the handle holds (persistent state + per-block QP scratch)
size_t ogb_sqp_bytes(void* h) {
    OgbDeviceSqp* q = (OgbDeviceSqp*)h;
    if (!q) return 0;
    return ((size_t)q->max_batch * q->T.S.state_doubles + (size_t)q->blocks * q->T.S.scratch_doubles) * sizeof(double);
}

// (re)start B instances: the next ogb_sqp_step treats c / vals as the values at the start points
int ogb_sqp_start(void* h, int B, void* stream) {
    OgbDeviceSqp* q = (OgbDeviceSqp*)h;
    if (!q || B <= 0 || B > q->max_batch) return ogb_set_error_message("ogb_sqp_start: bad handle or batch size");
    ogb_sqp_reset_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(q->T.S, q->state_d, B);
    OGS_CUDA(cudaGetLastError());
    q->launches += 1;
    return 0;
}

// one reverse-communication step of every instance.  x [B, nvars] in / out, c [B, m + 1] and vals [B, nnz] as
// written by ogb_eval_sparse / ogb_eval_exact at x; mode_h [B] (host, may be NULL) receives SLSQP's mode per
// instance after the step (1: evaluate again (line search), -1: evaluate again (gradients needed), else finished);
// with mode_h the call synchronises the stream.
int ogb_sqp_step(void* h, double* x, const double* c, const double* vals, int B, int32_t* mode_h, void* stream) {
    OgbDeviceSqp* q = (OgbDeviceSqp*)h;
    if (!q || !x || !c || !vals) return ogb_set_error_message("ogb_sqp_step: null argument");
    if (B <= 0 || B > q->max_batch) return ogb_set_error_message("ogb_sqp_step: bad batch size");
    cudaStream_t st = (cudaStream_t)stream;
    OGS_CUDA(cudaMemsetAsync(q->ticket_d, 0, sizeof(int), st));
    const int blocks = std::min(q->blocks, B);
    {
        OgsShape S = q->T.S;
        double *state = q->state_d, *scratch = q->scratch_d;
        int *ticket = q->ticket_d, *mode_d = q->mode_d, wrows = q->wrows, Bv = B, flags = q->flags;
        void* kargs[] = {&S, &x, &c, &vals, &state, &scratch, &Bv, &ticket, &mode_d, &wrows, &flags};
        OGS_CUDA(cudaLaunchKernel(q->fn, dim3((unsigned)blocks), dim3((unsigned)q->threads), kargs, q->smem, st));
    }
    OGS_CUDA(cudaGetLastError());
    q->launches += 1;
    if (mode_h) {
        OGS_CUDA(cudaMemcpyAsync(mode_h, q->mode_d, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
        OGS_CUDA(cudaStreamSynchronize(st));
    }
    return 0;
}

// the SLSQP scalars of every instance (host, [B, 24] doubles: f, f0, gs, h1, h2, h3, h4, t, t0, alpha, mode,
// iter, reset, line, inconsistent, nfev, njev, ...); synchronises the stream
int ogb_sqp_scalars(void* h, int B, double* sc_h, void* stream) {
    OgbDeviceSqp* q = (OgbDeviceSqp*)h;
    if (!q || !sc_h || B <= 0 || B > q->max_batch) return ogb_set_error_message("ogb_sqp_scalars: bad argument");
    const OgsShape& S = q->T.S;
    OGS_CUDA(cudaMemcpy2DAsync(sc_h, OGS_NSC * sizeof(double), q->state_d + S.o_sc, S.state_doubles * sizeof(double),
                               OGS_NSC * sizeof(double), (size_t)B, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    OGS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

long long ogb_sqp_launches(void* h) { return h ? ((OgbDeviceSqp*)h)->launches : 0; }

}  // extern "C"
