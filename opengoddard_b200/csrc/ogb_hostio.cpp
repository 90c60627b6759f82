// ogb_hostio.cpp -- the host-buffer entry point of libogb200.so (include/ogb200.h:
// ogb_host_session_*, ogb_host_eval_fd, ogb_host_expand).  Host-only C++ on top of the device
// ABI (ogb_dx_gemm / ogb_sweep / ogb_pack): it owns no arithmetic, only the transport.
//
// What it stands in for: a host caller of the reference gets c and the dense FD Jacobian as numpy
// arrays in host memory (reference OpenGoddard/optimize.py:711-715 through
// scipy/optimize/_slsqp_py.py:353-367).  Here B instances are served per call: the batch is cut
// into chunks, each chunk runs  H2D p -> K1 -> K2 -> K3 (pack) on the compute stream and its
// packed values + c come back on a second stream while the next chunk computes; a pool of host
// threads turns packed values into the dense column-major J the caller asked for, writing the
// zeros with non-temporal stores so host DRAM sees every byte once.
#include <cuda_runtime_api.h>
#include <immintrin.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ogb200.h"

extern "C" void ogb_set_error_text(const char* msg);     // ogb_kernels.cu (thread-local error text)

namespace {

int fail(const std::string& m) { ogb_set_error_text(m.c_str()); return -1; }

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

int usable_cores() {
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) return std::max(1, CPU_COUNT(&set));
    return std::max(1u, std::thread::hardware_concurrency());
}

// Scattered 8-byte stores into a matrix that is not in cache: every touched line costs a read-for-ownership
// miss, and a core only keeps a handful of them in flight on its own.  Prefetching the destination lines a
// hundred entries ahead ($OGB200_HOST_PREFETCH, default 128; 0 = off) raises the memory-level parallelism
// (Goddard-50 x 4096 on 16 threads: scatter 18.7 -> 15.1 ms, keep-zeros 18.7 -> 14.0 ms; tools/scatter_probe.py).
int scatter_prefetch_distance() {
    static const int d = [] {
        const char* v = getenv("OGB200_HOST_PREFETCH");
        const int k = v ? atoi(v) : 128;
        return k < 0 ? 0 : (k > 1024 ? 1024 : k);
    }();
    return d;
}

int host_isa() {
    // $OGB200_HOST_ISA = 0 / 1 / 2 picks SSE2 / AVX2 / AVX-512 stores.  Default AVX2: one core's
    // non-temporal write rate (~13 GB/s) is the limit either way and 512-bit stores measured slower.
    const char* f = getenv("OGB200_HOST_ISA");
    const int cap = f ? atoi(f) : 1;
    if (cap >= 2 && __builtin_cpu_supports("avx512f")) return 2;
    if (cap >= 1 && __builtin_cpu_supports("avx2")) return 1;
    return 0;
}

// ---------------------------------------------------------------- dense expansion of one instance
// dst[0, nM) = 0 except dst[lin[e]] = vals[e]; lin ascending.  The body is written one 64-byte line
// at a time with non-temporal stores (no read-for-ownership): runs of all-zero lines (~90 % of a
// Jacobian) come straight from a zero register, a line that holds values is assembled in a
// register-sized scratch first.  isa: 1 = AVX2 (32-byte stores), 0 = SSE2, 2 = AVX-512.
__attribute__((target("avx2"))) size_t zero_lines_avx2(double* dst, size_t pos, size_t stop) {
    const __m256d z = _mm256_setzero_pd();
    for (; pos + 8 <= stop; pos += 8) { _mm256_stream_pd(dst + pos, z); _mm256_stream_pd(dst + pos + 4, z); }
    return pos;
}
__attribute__((target("avx2"))) void put_line_avx2(double* dst, const double* line) {
    _mm256_stream_pd(dst, _mm256_load_pd(line));
    _mm256_stream_pd(dst + 4, _mm256_load_pd(line + 4));
}
__attribute__((target("avx512f"))) size_t zero_lines_avx512(double* dst, size_t pos, size_t stop) {
    const __m512d z = _mm512_setzero_pd();
    for (; pos + 8 <= stop; pos += 8) _mm512_stream_pd(dst + pos, z);
    return pos;
}
__attribute__((target("avx512f"))) void put_line_avx512(double* dst, const double* line) {
    _mm512_stream_pd(dst, _mm512_load_pd(line));
}
size_t zero_lines_sse2(double* dst, size_t pos, size_t stop) {
    const __m128d z = _mm_setzero_pd();
    for (; pos + 8 <= stop; pos += 8)
        for (int k = 0; k < 8; k += 2) _mm_stream_pd(dst + pos + k, z);
    return pos;
}
void put_line_sse2(double* dst, const double* line) {
    for (int k = 0; k < 8; k += 2) _mm_stream_pd(dst + k, _mm_load_pd(line + k));
}

void expand_dense(double* dst, size_t nM, const double* vals, const uint32_t* lin, int nnz, int isa) {
    size_t pos = 0;
    int e = 0;
    size_t head = ((64 - (reinterpret_cast<uintptr_t>(dst) & 63)) & 63) / sizeof(double);
    if (head > nM) head = nM;
    for (; pos < head; ++pos) dst[pos] = (e < nnz && lin[e] == pos) ? vals[e++] : 0.0;
    const size_t body_end = head + ((nM - head) & ~(size_t)7);       // whole 64-byte lines: [head, body_end)
    alignas(64) double line[8];
    while (pos < body_end) {
        // all-zero lines up to the line that holds the next value (or to the end of the body)
        size_t stop = body_end;
        if (e < nnz && lin[e] < body_end) stop = head + ((lin[e] - head) & ~(size_t)7);
        if (pos < stop)
            pos = isa == 1 ? zero_lines_avx2(dst, pos, stop) : isa == 2 ? zero_lines_avx512(dst, pos, stop)
                                                                        : zero_lines_sse2(dst, pos, stop);
        if (pos >= body_end) break;
        for (int k = 0; k < 8; ++k) line[k] = 0.0;                  // a line with values
        while (e < nnz && lin[e] < pos + 8) { line[lin[e] - pos] = vals[e]; ++e; }
        if (isa == 1) put_line_avx2(dst + pos, line);
        else if (isa == 2) put_line_avx512(dst + pos, line);
        else put_line_sse2(dst + pos, line);
        pos += 8;
    }
    for (; pos < nM; ++pos) dst[pos] = (e < nnz && lin[e] == pos) ? vals[e++] : 0.0;
    _mm_sfence();
}



void expand_keep(double* dst, const double* vals, const uint32_t* lin, int nnz) {
    const int pf = scatter_prefetch_distance();
    for (int e = 0; e < nnz; ++e) {
        if (pf && e + pf < nnz) __builtin_prefetch(dst + lin[e + pf], 1, 0);
        dst[lin[e]] = vals[e];
    }
}

// ---------------------------------------------------------------- worker pool
struct Job {
    const double* vals = nullptr;    // [B, nnz] packed (pinned staging), null in DMA mode
    const double* c_src = nullptr;   // [B, M] staging
    double* c_dst = nullptr;
    double* J_dst = nullptr;
    const uint32_t* lin = nullptr;
    int nnz = 0, M = 0, B = 0, chunk = 1, mode = 0;
    size_t nM = 0;
    int isa = 0;
    // mode OGB_HOST_J_SCATTER: packed entry e of instance b goes to C_dst[b][soff[e]] (soff >= 0) or
    // to g_dst[b][-1 - soff[e]] (the cost row), or nowhere (soff = INT64_MIN)
    double* const* C_dst = nullptr;
    double* const* g_dst = nullptr;
    const int64_t* soff = nullptr;
    int n = 0;
};

constexpr int kScatterMode = 100;       // internal: ogb_host_eval_fd_scatter
constexpr int64_t kNowhere = INT64_MIN;

void expand_scatter(double* C, double* g, const double* vals, const int64_t* soff, int nnz) {
    const int pf = scatter_prefetch_distance();
    for (int e = 0; e < nnz; ++e) {
        if (pf && e + pf < nnz) {
            const int64_t o2 = soff[e + pf];
            if (o2 >= 0) __builtin_prefetch(C + o2, 1, 0);
        }
        const int64_t o = soff[e];
        if (o >= 0) C[o] = vals[e];
        else if (o != kNowhere && g) g[-1 - o] = vals[e];
    }
}

class Pool {
public:
    explicit Pool(int n) {
        for (int t = 0; t < n; ++t) th_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; ++gen_; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    int size() const { return (int)th_.size(); }
    void start(const Job& j) {
        { std::lock_guard<std::mutex> g(m_); job_ = j; next_.store(0); ready_.store(0); running_ = size(); ++gen_; }
        cv_.notify_all();
    }
    void publish(int chunks_ready) { ready_.store(chunks_ready, std::memory_order_release); }
    void wait() {
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this] { return running_ == 0; });
    }
    static void run_one(const Job& j, int b) {
        if (j.c_dst) std::memcpy(j.c_dst + (size_t)b * j.M, j.c_src + (size_t)b * j.M, (size_t)j.M * sizeof(double));
        if (!j.vals) return;
        const double* v = j.vals + (size_t)b * j.nnz;
        if (j.mode == kScatterMode) expand_scatter(j.C_dst[b], j.g_dst ? j.g_dst[b] : nullptr, v, j.soff, j.nnz);
        else if (j.mode == OGB_HOST_J_DENSE) expand_dense(j.J_dst + (size_t)b * j.nM, j.nM, v, j.lin, j.nnz, j.isa);
        else if (j.mode == OGB_HOST_J_KEEP_ZEROS) expand_keep(j.J_dst + (size_t)b * j.nM, v, j.lin, j.nnz);
        else std::memcpy(j.J_dst + (size_t)b * j.nnz, v, (size_t)j.nnz * sizeof(double));
    }

private:
    void loop() {
        unsigned long seen = 0;
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                j = job_;
            }
            for (;;) {
                const int b = next_.fetch_add(1);
                if (b >= j.B) break;
                const int k = b / j.chunk;
                int spins = 0;
                while (ready_.load(std::memory_order_acquire) <= k) {
                    if (++spins < 200) _mm_pause(); else { std::this_thread::yield(); spins = 0; }
                }
                run_one(j, b);
            }
            { std::lock_guard<std::mutex> g(m_); if (--running_ == 0) done_.notify_all(); }
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    Job job_;
    std::atomic<int> next_{0}, ready_{0};
    int running_ = 0;
    unsigned long gen_ = 0;
    bool stop_ = false;
};

#define HCUDA(call)                                                                          \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct Session {
    void* prob = nullptr;
    ogb_problem_info info{};
    int maxB = 0, chunk = 0, nchunks = 0, device = 0;
    int nnz = 0;
    size_t nM = 0;
    std::vector<uint32_t> lin;
    int isa = 0;
    // device
    double *p_d = nullptr, *lb_d = nullptr, *ub_d = nullptr, *c_d = nullptr, *vals_d = nullptr;
    double *DX_d = nullptr, *Jfull_d = nullptr;
    size_t work_bytes = 0;
    int exact = 0;                  // 1: the session's Jacobians come from ogb_eval_exact (opt-in)
    std::vector<int64_t> soff;      // scatter offsets of the last ogb_host_eval_fd_scatter layout
    int soff_ld = -1, soff_mrows = -1;
    // pinned staging
    double *p_s = nullptr, *c_s = nullptr, *vals_s = nullptr, *b_s = nullptr;
    cudaStream_t s_compute = nullptr, s_copy = nullptr;
    std::vector<cudaEvent_t> ev_done, ev_copied;
    Pool* pool = nullptr;
    ogb_host_stats st{};

    ~Session() {
        delete pool;
        cudaSetDevice(device);
        for (double* d : {p_d, lb_d, ub_d, c_d, vals_d, DX_d, Jfull_d}) cudaFree(d);
        for (double* h : {p_s, c_s, vals_s, b_s}) cudaFreeHost(h);
        for (cudaEvent_t e : ev_done) cudaEventDestroy(e);
        for (cudaEvent_t e : ev_copied) cudaEventDestroy(e);
        if (s_compute) cudaStreamDestroy(s_compute);
        if (s_copy) cudaStreamDestroy(s_copy);
    }
};

bool is_pinned(const void* ptr) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int session_init(Session* S, void* prob, int max_batch, int chunk, int threads) {
    S->prob = prob;
    if (ogb_problem_info_get(prob, &S->info) != 0) return -1;
    HCUDA(cudaGetDevice(&S->device));
    const int n = S->info.nvars, M = S->info.nrows;
    S->nM = (size_t)n * M;
    S->nnz = ogb_jac_pattern(prob, nullptr, 0);
    if (S->nnz < 0) return -1;
    S->lin.resize(std::max(1, S->nnz));
    if (ogb_jac_pattern(prob, S->lin.data(), S->nnz) < 0) return -1;
    S->maxB = max_batch;
    if (chunk <= 0) {
        // default: ~12 chunks per full batch (each chunk's device->host copy overlaps the next chunk's
        // kernels and the host threads' work on the previous one), at least 32 MB of packed values each
        size_t per = std::max<size_t>(1, (size_t)S->nnz * sizeof(double));
        chunk = (int)std::max<size_t>(16, (32u << 20) / per);
        chunk = std::max(chunk, (max_batch + 11) / 12);
    }
    S->chunk = std::min(chunk, max_batch);
    S->nchunks = (max_batch + S->chunk - 1) / S->chunk;
    S->isa = host_isa();
    const size_t B = (size_t)max_batch;
    HCUDA(cudaMalloc((void**)&S->p_d, B * n * 8));
    HCUDA(cudaMalloc((void**)&S->lb_d, (size_t)n * 8));
    HCUDA(cudaMalloc((void**)&S->ub_d, (size_t)n * 8));
    HCUDA(cudaMalloc((void**)&S->c_d, B * M * 8));
    HCUDA(cudaMalloc((void**)&S->vals_d, std::max<size_t>(1, B * S->nnz) * 8));
    S->work_bytes = std::max<size_t>(256, ogb_workspace_bytes(prob, S->chunk));
    HCUDA(cudaMalloc((void**)&S->DX_d, S->work_bytes));
    HCUDA(cudaMallocHost((void**)&S->p_s, B * n * 8));
    HCUDA(cudaMallocHost((void**)&S->c_s, B * M * 8));
    HCUDA(cudaMallocHost((void**)&S->vals_s, std::max<size_t>(1, B * S->nnz) * 8));
    HCUDA(cudaMallocHost((void**)&S->b_s, (size_t)2 * n * 8));
    HCUDA(cudaStreamCreateWithFlags(&S->s_compute, cudaStreamNonBlocking));
    HCUDA(cudaStreamCreateWithFlags(&S->s_copy, cudaStreamNonBlocking));
    S->ev_done.resize(S->nchunks);
    S->ev_copied.resize(S->nchunks);
    for (int k = 0; k < S->nchunks; ++k) {
        HCUDA(cudaEventCreateWithFlags(&S->ev_done[k], cudaEventDisableTiming));
        HCUDA(cudaEventCreateWithFlags(&S->ev_copied[k], cudaEventDisableTiming | cudaEventBlockingSync));
    }
    if (threads <= 0) threads = usable_cores();
    S->pool = new Pool(std::max(1, std::min(threads, 256)));
    return 0;
}

}  // namespace

extern "C" {

void* ogb_host_session_create(void* prob, int max_batch, int chunk, int threads) {
    if (!prob || max_batch <= 0) { fail("ogb_host_session_create: bad argument"); return nullptr; }
    Session* S = new Session();
    if (session_init(S, prob, max_batch, chunk, threads) != 0) { delete S; return nullptr; }
    return S;
}

void ogb_host_session_destroy(void* h) { delete (Session*)h; }

int ogb_host_session_set_option(void* h, int key, int value) {
    Session* S = (Session*)h;
    if (!S) return fail("ogb_host_session_set_option: null session");
    if (key == OGB_HOST_OPT_EXACT) { S->exact = value != 0; return 0; }
    return fail("ogb_host_session_set_option: unknown key");
}

int ogb_host_session_stats(void* h, ogb_host_stats* out) {
    Session* S = (Session*)h;
    if (!S || !out) return fail("ogb_host_session_stats: null argument");
    *out = S->st;
    return 0;
}

// Shared body of ogb_host_eval_fd / ogb_host_eval_fd_scatter.  `job` arrives with its destination
// fields (mode, J_dst or C_dst / g_dst / soff) filled in.
static int host_eval_core(Session* S, const double* p_h, const double* lb_h, const double* ub_h, double abs_step,
                          int B, double* c_h, double* J_h, Job job) {
    const double t0 = now_ms();
    HCUDA(cudaSetDevice(S->device));
    const int n = S->info.nvars, M = S->info.nrows, CH = S->chunk;
    const int nch = (B + CH - 1) / CH;
    const int mode = job.mode;
    const bool dma = mode == OGB_HOST_J_DMA;
    if (dma) {
        if (!S->Jfull_d) HCUDA(cudaMalloc((void**)&S->Jfull_d, (size_t)S->maxB * S->nM * 8));
        const size_t need = ogb_workspace_bytes(S->prob, S->chunk);      // (the dense path may want more scratch)
        if (need > S->work_bytes) {
            cudaFree(S->DX_d);
            S->DX_d = nullptr;
            HCUDA(cudaMalloc((void**)&S->DX_d, need));
            S->work_bytes = need;
        }
    }
    ogb_problem_info inf0{};
    ogb_problem_info_get(S->prob, &inf0);
    int64_t h2d = 0, d2h = 0;

    std::memcpy(S->b_s, lb_h, (size_t)n * 8);
    std::memcpy(S->b_s + n, ub_h, (size_t)n * 8);
    HCUDA(cudaMemcpyAsync(S->lb_d, S->b_s, (size_t)n * 8, cudaMemcpyHostToDevice, S->s_compute));
    HCUDA(cudaMemcpyAsync(S->ub_d, S->b_s + n, (size_t)n * 8, cudaMemcpyHostToDevice, S->s_compute));
    h2d += 2 * (int64_t)n * 8;
    const bool p_pinned = is_pinned(p_h);
    if (p_pinned) HCUDA(cudaMemcpyAsync(S->p_d, p_h, (size_t)B * n * 8, cudaMemcpyHostToDevice, S->s_compute));
    h2d += (int64_t)B * n * 8;

    job.vals = dma ? nullptr : S->vals_s;
    job.c_src = S->c_s; job.c_dst = c_h; job.lin = S->lin.data();
    job.nnz = S->nnz; job.M = M; job.B = B; job.chunk = CH; job.nM = S->nM; job.isa = S->isa; job.n = n;
    S->pool->start(job);

    int rc = 0;
    for (int k = 0; k < nch && !rc; ++k) {
        const int b0 = k * CH, nb = std::min(CH, B - b0);
        double* pk = S->p_d + (size_t)b0 * n;
        if (!p_pinned) {
            std::memcpy(S->p_s + (size_t)b0 * n, p_h + (size_t)b0 * n, (size_t)nb * n * 8);
            if (cudaMemcpyAsync(pk, S->p_s + (size_t)b0 * n, (size_t)nb * n * 8, cudaMemcpyHostToDevice, S->s_compute) != cudaSuccess) { rc = fail("H2D of p failed"); break; }
        }
        double* ck = S->c_d + (size_t)b0 * M;
        if (dma) {          // the plain dense transport: the dense J in HBM, one device -> host copy
            rc = ogb_eval_fd(S->prob, pk, S->lb_d, S->ub_d, abs_step, nb, ck, S->Jfull_d + (size_t)b0 * S->nM, S->DX_d, S->s_compute);
        } else {            // K1 + the sweep kernel with packed output: no dense J anywhere on the device
            rc = S->exact
                ? ogb_eval_exact(S->prob, pk, S->lb_d, S->ub_d, nb, ck, S->vals_d + (size_t)b0 * S->nnz, S->DX_d, S->s_compute)
                : ogb_eval_sparse(S->prob, pk, S->lb_d, S->ub_d, abs_step, nb, ck, S->vals_d + (size_t)b0 * S->nnz, S->DX_d, S->s_compute);
        }
        if (rc) break;
        cudaEventRecord(S->ev_done[k], S->s_compute);
        cudaStreamWaitEvent(S->s_copy, S->ev_done[k], 0);
        cudaMemcpyAsync(S->c_s + (size_t)b0 * M, ck, (size_t)nb * M * 8, cudaMemcpyDeviceToHost, S->s_copy);
        d2h += (int64_t)nb * M * 8;
        if (dma) {
            cudaMemcpyAsync(J_h + (size_t)b0 * S->nM, S->Jfull_d + (size_t)b0 * S->nM, (size_t)nb * S->nM * 8, cudaMemcpyDeviceToHost, S->s_copy);
            d2h += (int64_t)nb * S->nM * 8;
        } else {
            cudaMemcpyAsync(S->vals_s + (size_t)b0 * S->nnz, S->vals_d + (size_t)b0 * S->nnz,
                            (size_t)nb * S->nnz * 8, cudaMemcpyDeviceToHost, S->s_copy);
            d2h += (int64_t)nb * S->nnz * 8;
        }
        if (cudaEventRecord(S->ev_copied[k], S->s_copy) != cudaSuccess) rc = fail("cudaEventRecord failed");
    }
    double t_first = 0.0;
    cudaError_t ce = cudaSuccess;
    if (!rc) {
        for (int k = 0; k < nch; ++k) {
            ce = cudaEventSynchronize(S->ev_copied[k]);
            if (ce != cudaSuccess) break;
            if (k == 0) t_first = now_ms() - t0;
            S->pool->publish(k + 1);
        }
    }
    S->pool->publish(nch + 1);                 // on error: release the workers anyway
    S->pool->wait();
    if (rc) { cudaStreamSynchronize(S->s_compute); cudaStreamSynchronize(S->s_copy); return rc; }
    if (ce != cudaSuccess) return fail(std::string("ogb_host_eval_fd: ") + cudaGetErrorString(ce));
    ogb_problem_info inf1{};
    ogb_problem_info_get(S->prob, &inf1);
    S->st.h2d_bytes = h2d; S->st.d2h_bytes = d2h; S->st.launches = (int32_t)(inf1.launches - inf0.launches); S->st.nnz = S->nnz;
    S->st.chunk = CH; S->st.threads = S->pool->size(); S->st.nchunks = nch;
    S->st.ms_total = now_ms() - t0; S->st.ms_first_chunk = t_first;
    return 0;
}

int ogb_host_eval_fd(void* h, const double* p_h, const double* lb_h, const double* ub_h, double abs_step,
                     int B, double* c_h, double* J_h, int mode) {
    Session* S = (Session*)h;
    if (S && B == 0) return 0;
    if (!S || !p_h || !lb_h || !ub_h || !c_h || !J_h) return fail("ogb_host_eval_fd: null argument");
    if (B < 0 || B > S->maxB) return fail("ogb_host_eval_fd: batch larger than the session's max_batch");
    if (mode < OGB_HOST_J_DENSE || mode > OGB_HOST_J_DMA) return fail("ogb_host_eval_fd: unknown mode");
    if (!(abs_step > 0.0)) return fail("ogb_host_eval_fd: abs_step must be positive");
    Job job;
    job.mode = mode;
    job.J_dst = J_h;
    return host_eval_core(S, p_h, lb_h, ub_h, abs_step, B, c_h, J_h, job);
}

int ogb_host_eval_fd_scatter(void* h, const double* p_h, const double* lb_h, const double* ub_h, double abs_step,
                             int B, double* c_h, double* const* C_h, int ld, int mrows, double* const* g_h) {
    Session* S = (Session*)h;
    if (S && B == 0) return 0;
    if (!S || !p_h || !lb_h || !ub_h || !c_h || !C_h) return fail("ogb_host_eval_fd_scatter: null argument");
    if (B < 0 || B > S->maxB) return fail("ogb_host_eval_fd_scatter: batch larger than the session's max_batch");
    if (!(abs_step > 0.0)) return fail("ogb_host_eval_fd_scatter: abs_step must be positive");
    const int M = S->info.nrows;
    if (mrows < 0 || mrows > M || ld < std::max(1, mrows)) return fail("ogb_host_eval_fd_scatter: need 0 <= mrows <= nrows and ld >= max(1, mrows)");
    if (S->soff_ld != ld || S->soff_mrows != mrows) {      // destination offset of every packed entry, once per layout
        S->soff.assign((size_t)std::max(1, S->nnz), kNowhere);
        for (int e = 0; e < S->nnz; ++e) {
            const uint32_t l = S->lin[e];
            const int64_t j = l / (uint32_t)M, r = l - (uint32_t)j * (uint32_t)M;
            if (r < mrows) S->soff[e] = j * (int64_t)ld + r;
            else if (r == M - 1) S->soff[e] = -1 - j;
        }
        S->soff_ld = ld; S->soff_mrows = mrows;
    }
    Job job;
    job.mode = kScatterMode;
    job.C_dst = C_h; job.g_dst = g_h; job.soff = S->soff.data();
    return host_eval_core(S, p_h, lb_h, ub_h, abs_step, B, c_h, nullptr, job);
}

int ogb_host_expand(const double* vals_h, const uint32_t* lin_h, int nnz, size_t nM, int B, double* J_h,
                    int mode, int threads) {
    if (!vals_h || !lin_h || !J_h || nnz < 0 || B < 0) return fail("ogb_host_expand: bad argument");
    if (mode != OGB_HOST_J_DENSE && mode != OGB_HOST_J_KEEP_ZEROS) return fail("ogb_host_expand: mode must be dense or keep-zeros");
    for (int e = 0; e < nnz; ++e)
        if (lin_h[e] >= nM || (e && lin_h[e] <= lin_h[e - 1])) return fail("ogb_host_expand: pattern must be ascending and < nM");
    if (threads <= 0) threads = usable_cores();
    threads = std::max(1, std::min(threads, std::max(1, B)));
    Job j;
    j.vals = vals_h; j.J_dst = J_h; j.lin = lin_h; j.nnz = nnz; j.B = B; j.chunk = std::max(1, B); j.mode = mode;
    j.nM = nM; j.isa = host_isa();
    std::atomic<int> next{0};
    auto work = [&] { for (int b; (b = next.fetch_add(1)) < B;) Pool::run_one(j, b); };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return 0;
}

}  // extern "C"
