// ogb_sqp.h -- one SLSQP iteration per problem instance, written for a cooperating group of threads.
//
// SURVEY.md section 8f row 1 ("device-side batched QP"): after the FD Jacobian moved to the GPU, SLSQP's
// own core is what a multi-start spends its time in (scipy/optimize/_slsqp_py.py:524-555 drives it for the
// reference's Problem.solve, /root/reference/OpenGoddard/optimize.py:738-755).  This file restates Kraft's
// SLSQP -- the outer iteration `slsqpb` and its least-squares core LSQ -> LSEI -> LSI -> LDP -> NNLS with
// Householder (H12) and Givens transformations (D. Kraft, DFVLR-FB 88-28, 1988; Lawson & Hanson, "Solving
// Least Squares Problems", 1974) -- so that ONE thread block advances ONE instance, every matrix in global
// memory (L2-resident while the block works on it), all loops over rows / columns spread over the block:
//
//   * rows of C, E, G are contiguous: a Householder transformation from the right is a warp-per-row
//     dot + axpy (coalesced);
//   * the NNLS matrix is stored constraint-minor ((l + 1) x mg): one thread per constraint for the dual
//     vector, the Householder application and the Givens rotations (coalesced);
//   * the short sequential recurrences (triangular solves, the L D L' rank-one updates) run as block-wide
//     axpy steps.
//
// Everything is a template over a "group context" Cx (tid / nthr / warp / lane, sync, block and warp
// reductions): the CUDA kernel (ogb_sqp.cu) instantiates it with a thread block, tests/emu with ONE
// serial thread, so the container without a GPU can check the very same code against the numpy
// restatement in the oracle directory.  No part of this is a CPU fallback: the product only runs the kernel.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __CUDACC__
#define OGS_HD __host__ __device__ __forceinline__
#define OGS_FN __host__ __device__
#else
#define OGS_HD inline
#define OGS_FN
#endif

// fused multiply-add in the hot loops of the QP (one FP64 pipe slot instead of two); this solver is pinned to the
// restatement by tolerance, not bitwise, so contraction is allowed here (the sweep kernel stays -fmad=false)
#define OGS_FMA(a, b, c) fma((a), (b), (c))
#define OGS_EPS 2.220446049250313e-16
#define OGS_ALFMIN 0.1
#define OGS_INF (1.0 / 0.0)
#define OGS_MODE_START 100          // state code of an instance that has not taken its first step (0 is SLSQP's success)

// ---- problem shape shared by every instance of a batch
struct OgsShape {
    int n, m, meq, mineq, M, nnz;      // M = m + 1: the cost is the last row of c / of the Jacobian
    int nlo, nhi;                      // variables with a finite lower / upper bound
    int n1, mg, lmax;                  // n + 1;  mineq + nlo + nhi + 2;  n1 - meq
    const int* colptr;                 // [n + 1] packed Jacobian, by variable (ogb_jac_pattern)
    const int* prow;                   // [nnz]   row of every packed entry
    const int* rowptr;                 // [M + 1] the same entries by row
    const int* rcol;                   // [nnz]   variable of the entry
    const int* rpos;                   // [nnz]   position in the packed values
    const int* blo;                    // [nlo]
    const int* bhi;                    // [nhi]
    const double* xl;                  // [n] (-inf = none)
    const double* xu;                  // [n] (+inf = none)
    double acc;
    int itermax;
    // per-instance persistent state, offsets in doubles
    size_t o_x0, o_s, o_g, o_mu, o_r, o_v, o_u, o_w, o_lt, o_dg, o_sc, state_doubles;
    // per-block scratch, offsets in doubles
    size_t o_E, o_f, o_C, o_d, o_ups, o_G, o_h, o_xq, o_wq, o_E2T, o_GT, o_An, o_bn, o_zn, o_xn, o_wn, o_ucol,
        o_f2, o_fres, o_binv, o_idx, scratch_doubles;
};

enum { OGS_F = 0, OGS_F0, OGS_GS, OGS_H1, OGS_H2, OGS_H3, OGS_H4, OGS_T, OGS_T0, OGS_ALPHA, OGS_MODE, OGS_ITER,
       OGS_RESET, OGS_LINE, OGS_BADLIN, OGS_NFEV, OGS_NJEV,
       OGS_CLK_BUILD = 18, OGS_CLK_LSEI, OGS_CLK_LSI, OGS_CLK_NNLS, OGS_CLK_FINISH, OGS_CLK_BFGS,   // SM cycles per phase
       OGS_NSC = 24 };

static inline size_t ogs_al(size_t x) { return (x + 3) & ~(size_t)3; }

// fills the derived sizes and the two layouts (host side)
static inline void ogs_layout(OgsShape& S) {
    S.mineq = S.m - S.meq;
    S.M = S.m + 1;
    S.n1 = S.n + 1;
    S.mg = S.mineq + S.nlo + S.nhi + 2;
    S.lmax = S.n1 - S.meq;
    size_t o = 0;
    auto take = [&](size_t cnt) { size_t at = o; o += ogs_al(cnt); return at; };
    S.o_x0 = take(S.n); S.o_s = take(S.n1); S.o_g = take(S.n1); S.o_mu = take(S.m); S.o_r = take(S.m);
    S.o_v = take(S.n); S.o_u = take(S.n); S.o_w = take(S.n); S.o_lt = take((size_t)S.n * S.n); S.o_dg = take(S.n);
    S.o_sc = take(OGS_NSC);
    S.state_doubles = o;
    o = 0;
    const size_t n1 = S.n1, mg = S.mg, l1 = S.lmax + 1;
    S.o_E = take(n1 * n1); S.o_f = take(n1); S.o_C = take((size_t)S.meq * n1); S.o_d = take(S.meq);
    S.o_ups = take(S.meq); S.o_G = take(mg * n1); S.o_h = take(mg); S.o_xq = take(n1); S.o_wq = take(S.meq + mg);
    S.o_E2T = take((size_t)S.lmax * n1); S.o_GT = take(l1 * mg); S.o_An = take(l1 * mg); S.o_bn = take(l1);
    S.o_zn = take(l1); S.o_xn = take(mg); S.o_wn = take(mg); S.o_ucol = take(l1);
    S.o_f2 = take(n1); S.o_fres = take(n1); S.o_binv = take(S.meq);
    S.o_idx = take((mg + 1) / 2 + 1);
    S.scratch_doubles = o;
}

// ---- the serial group (tests/emu): one thread, one lane
struct OgsSerial {
    int tid = 0, nthr = 1, warp = 0, nwarps = 1, lane = 0, wsize = 1;
    OGS_HD void sync() {}
    OGS_HD double sum(double v) { return v; }
    OGS_HD double max(double v) { return v; }
    OGS_HD double wsum(double v) { return v; }
    OGS_HD void argbest(double& v, int& i, bool) {}
    OGS_HD long long clock() { return 0; }
    double* wb = nullptr;              // one row buffer (n + 1 doubles) per warp
    int wrows = 1, wstride = 0;        // rows a warp keeps in its buffer at once, doubles per row
    OGS_HD double* wbuf() { return wb; }
    OGS_HD void wsync() {}
    OGS_HD double wmax(double v) { return v; }
    static constexpr bool kLanes32 = false;   // (the register-resident reflector path needs 32 lanes)
    int flags = 0;
};

// ---------------------------------------------------------------- Householder (Lawson & Hanson H12)
// construct on u[p], u[l1 .. len) (stride `st`); returns `up`, sets ident when the transformation is I
template <class Cx>
OGS_FN double ogs_h12_construct(Cx& cx, double* u, size_t st, int p, int l1, int len, bool& ident) {
    ident = true;
    if (!(p < l1 && l1 < len)) return 0.0;
    double cl = 0.0;
    for (int k = l1 + cx.tid; k < len; k += cx.nthr) cl = fmax(cl, fabs(u[k * st]));
    cl = cx.max(cl);
    const double upiv = u[p * st];
    cl = fmax(cl, fabs(upiv));
    if (cl <= 0.0) return 0.0;
    const double clinv = 1.0 / cl;
    double sm = 0.0;
    for (int k = l1 + cx.tid; k < len; k += cx.nthr) { const double t = u[k * st] * clinv; sm += t * t; }
    sm = cx.sum(sm);
    sm += (upiv * clinv) * (upiv * clinv);
    cl = cl * sqrt(sm);
    if (upiv > 0.0) cl = -cl;
    cx.sync();
    if (cx.tid == 0) u[p * st] = cl;
    cx.sync();
    ident = false;
    return upiv - cl;
}

// apply to `nrows` contiguous rows of Cm (leading dimension ldc): one warp per row
template <class Cx>
OGS_FN void ogs_h12_rows(Cx& cx, const double* u, int p, int l1, int len, double up, bool ident, double* Cm, size_t ldc,
                         int nrows) {
    if (ident || nrows <= 0) return;
    double b = up * u[p];
    if (!(b < 0.0)) return;
    b = 1.0 / b;
    for (int r = cx.warp; r < nrows; r += cx.nwarps) {
        double* row = Cm + (size_t)r * ldc;
        double sm = 0.0;
        for (int k = l1 + cx.lane; k < len; k += cx.wsize) sm = OGS_FMA(row[k], u[k], sm);
        const double rp = row[p];
        sm = cx.wsum(sm);
        cx.wsync();
        sm += rp * up;
        if (sm != 0.0) {
            sm *= b;
            if (cx.lane == 0) row[p] = rp + sm * up;
            for (int k = l1 + cx.lane; k < len; k += cx.wsize) row[k] = OGS_FMA(sm, u[k], row[k]);
        }
    }
    cx.sync();
}

// apply to ONE contiguous vector with the whole group
template <class Cx>
OGS_FN void ogs_h12_vec(Cx& cx, const double* u, size_t st, int p, int l1, int len, double up, bool ident, double* c) {
    if (ident) return;
    double b = up * u[p * st];
    if (!(b < 0.0)) return;
    b = 1.0 / b;
    double sm = 0.0;
    for (int k = l1 + cx.tid; k < len; k += cx.nthr) sm += c[k] * u[k * st];
    const double cp = c[p];
    sm = cx.sum(sm);
    sm += cp * up;
    if (sm != 0.0) {
        sm *= b;
        if (cx.tid == 0) c[p] = cp + sm * up;
        for (int k = l1 + cx.tid; k < len; k += cx.nthr) c[k] += sm * u[k * st];
    }
    cx.sync();
}

// apply the reflection defined by u (pivot p, zeroing p + 1 .. len; binv = 1 / (up u[p]), 0 = identity) to R rows
// at once: rows[r] = base + r * stride (a warp's buffer in shared memory, or rows of a matrix in global memory).
// The R dot products and updates are independent, so their loads and shuffles overlap.
template <int R, class Cx>
OGS_FN void ogs_reflect_rows(Cx& cx, double* base, size_t stride, const double* u, int p, int len, double up, double binv) {
    if (binv == 0.0) return;
    double sm[R], rp[R];
#pragma unroll
    for (int r = 0; r < R; ++r) sm[r] = 0.0;
    for (int k = p + 1 + cx.lane; k < len; k += cx.wsize) {
        const double uk = u[k];
#pragma unroll
        for (int r = 0; r < R; ++r) sm[r] = OGS_FMA(base[r * stride + k], uk, sm[r]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) rp[r] = base[r * stride + p];
#pragma unroll
    for (int r = 0; r < R; ++r) sm[r] = cx.wsum(sm[r]);
    cx.wsync();                                   // (every lane has read its operands before any lane writes)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        sm[r] += rp[r] * up;
        if (sm[r] != 0.0) {
            sm[r] *= binv;
            if (cx.lane == 0) base[r * stride + p] = rp[r] + sm[r] * up;
        }
    }
    for (int k = p + 1 + cx.lane; k < len; k += cx.wsize) {
        const double uk = u[k];
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (sm[r] != 0.0) base[r * stride + k] = OGS_FMA(sm[r], uk, base[r * stride + k]);
    }
    cx.wsync();
}

// the rows [0, nrows) of a matrix in global memory (leading dimension ldc) through ONE reflection, 4 rows per warp at once
template <class Cx>
OGS_FN void ogs_reflect_matrix(Cx& cx, double* Cm, size_t ldc, int nrows, const double* u, int p, int len, double up, double binv) {
    if (binv == 0.0 || nrows <= 0) return;
    for (int r0 = cx.warp * 4; r0 < nrows; r0 += cx.nwarps * 4) {
        double* base = Cm + (size_t)r0 * ldc;
        const int cnt = nrows - r0;
        if (cnt >= 4) ogs_reflect_rows<4>(cx, base, ldc, u, p, len, up, binv);
        else if (cnt == 3) ogs_reflect_rows<3>(cx, base, ldc, u, p, len, up, binv);
        else if (cnt == 2) ogs_reflect_rows<2>(cx, base, ldc, u, p, len, up, binv);
        else ogs_reflect_rows<1>(cx, base, ldc, u, p, len, up, binv);
    }
}

// `cnt` (<= wrows) rows of a warp's buffer through the reflections p0 .. p1 - 1 defined by the rows of C
template <int R, class Cx>
OGS_FN void ogs_reflect_buffer(Cx& cx, double* rb, size_t stride, const double* C, size_t ld, const double* ups,
                               const double* binv, int p0, int p1, int len) {
    for (int p = p0; p < p1; ++p) ogs_reflect_rows<R>(cx, rb, stride, C + (size_t)p * ld, p, len, ups[p], binv[p]);
}

// The same for a 32-lane warp and at most 32 * NK columns: lane owns columns lane + 32 t.  The reflector of step
// p + 1 (and its up / binv) is fetched into registers while step p is being applied, so the L2 latency of the
// rows of C is hidden behind the arithmetic instead of being paid once per reflection.
template <int R, int NK, class Cx>
OGS_FN void ogs_reflect_buffer_reg(Cx& cx, double* rb, size_t stride, const double* C, size_t ld, const double* ups,
                                   const double* binv, int p0, int p1, int len) {
    if (p1 <= p0) return;
    double un[NK], upn, bvn;
    {
        const double* u = C + (size_t)p0 * ld;
#pragma unroll
        for (int t = 0; t < NK; ++t) { const int k = cx.lane + 32 * t; un[t] = (k > p0 && k < len) ? u[k] : 0.0; }
        upn = ups[p0]; bvn = binv[p0];
    }
    for (int p = p0; p < p1; ++p) {
        double uc[NK];
#pragma unroll
        for (int t = 0; t < NK; ++t) uc[t] = un[t];
        const double up = upn, bv = bvn;
        if (p + 1 < p1) {
            const double* u = C + (size_t)(p + 1) * ld;
#pragma unroll
            for (int t = 0; t < NK; ++t) { const int k = cx.lane + 32 * t; un[t] = (k > p + 1 && k < len) ? u[k] : 0.0; }
            upn = ups[p + 1]; bvn = binv[p + 1];
        }
        if (bv == 0.0) continue;
        double sm[R], rp[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double a = 0.0;
#pragma unroll
            for (int t = 0; t < NK; ++t) { const int k = cx.lane + 32 * t; if (k < len) a = OGS_FMA(rb[r * stride + k], uc[t], a); }
            sm[r] = a;
            rp[r] = rb[r * stride + p];
        }
#pragma unroll
        for (int r = 0; r < R; ++r) sm[r] = cx.wsum(sm[r]);
        cx.wsync();                               // (every lane has read its operands before any lane writes)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            sm[r] += rp[r] * up;
            if (sm[r] != 0.0) {
                sm[r] *= bv;
#pragma unroll
                for (int t = 0; t < NK; ++t) { const int k = cx.lane + 32 * t; if (k < len && uc[t] != 0.0) rb[r * stride + k] = OGS_FMA(sm[r], uc[t], rb[r * stride + k]); }
                if (cx.lane == 0) rb[r * stride + p] = rp[r] + sm[r] * up;
            }
        }
        cx.wsync();
    }
}

// ---------------------------------------------------------------- NNLS (Lawson & Hanson, chapter 23)
// min ||A x - b||, x >= 0;  A is (mrows x ncol) with leading dimension lda (constraint-minor), overwritten.
// Returns mode (1 / 3); x, w (dual), rnorm.
template <class Cx>
OGS_FN int ogs_nnls(Cx& cx, double* A, size_t lda, int mrows, int ncol, double* b, double* x, double* w, double* z,
                    double* ucol, int* idx, double& rnorm) {
    const double factor = 1.0e-2;
    for (int j = cx.tid; j < ncol; j += cx.nthr) { idx[j] = j; x[j] = 0.0; w[j] = 0.0; }
    cx.sync();
    int iz1 = 0, nsetp = 0, npp1 = 0, iter = 0, mode = 1;
    const int iz2 = ncol - 1, itmax = 3 * ncol;
    while (true) {
        if (iz1 > iz2 || nsetp >= mrows) break;
        for (int iz = iz1 + cx.tid; iz <= iz2; iz += cx.nthr) {
            const int j = idx[iz];
            double sm = 0.0;
            for (int i = npp1; i < mrows; ++i) sm += A[i * lda + j] * b[i];
            w[j] = sm;
        }
        cx.sync();
        bool found = false, ident = true;
        int j = -1, izs = -1;
        double up = 0.0;
        while (true) {
            double best = 0.0;
            int bi = -1;
            for (int iz = iz1 + cx.tid; iz <= iz2; iz += cx.nthr) {
                const double wj = w[idx[iz]];
                if (wj > best) { best = wj; bi = iz; }
            }
            cx.argbest(best, bi, true);                    // the largest w, first position on ties
            if (!(best > 0.0) || bi < 0) break;
            izs = bi;
            j = idx[izs];
            const double asave = A[npp1 * lda + j];
            up = ogs_h12_construct(cx, A + j, lda, npp1, npp1 + 1, mrows, ident);
            double un = 0.0;
            for (int i = cx.tid; i < nsetp; i += cx.nthr) un += A[i * lda + j] * A[i * lda + j];
            const double unorm = sqrt(cx.sum(un));
            const double apiv = A[npp1 * lda + j];
            const double t = factor * fabs(apiv);
            bool ok = false;
            if ((unorm + t) - unorm > 0.0) {
                for (int i = cx.tid; i < mrows; i += cx.nthr) z[i] = b[i];
                cx.sync();
                ogs_h12_vec(cx, A + j, lda, npp1, npp1 + 1, mrows, up, ident, z);
                if (z[npp1] / apiv > 0.0) ok = true;
            }
            if (ok) { found = true; break; }
            cx.sync();
            if (cx.tid == 0) { A[npp1 * lda + j] = asave; w[j] = 0.0; }
            cx.sync();
        }
        if (!found) break;
        // ---- step 5: column j enters the set P
        for (int i = cx.tid; i < mrows; i += cx.nthr) b[i] = z[i];
        const int jfirst = idx[iz1];
        cx.sync();
        if (cx.tid == 0) { idx[izs] = jfirst; idx[iz1] = j; }
        iz1 += 1;
        nsetp = npp1 + 1;
        npp1 += 1;
        for (int i = nsetp - 1 + cx.tid; i < mrows; i += cx.nthr) ucol[i] = A[i * lda + j];
        cx.sync();
        if (!ident && iz1 <= iz2) {
            double bb = up * ucol[nsetp - 1];
            if (bb < 0.0) {
                bb = 1.0 / bb;
                for (int iz = iz1 + cx.tid; iz <= iz2; iz += cx.nthr) {
                    const int jj = idx[iz];
                    const double cp = A[(nsetp - 1) * lda + jj];
                    double sm = 0.0;
                    for (int i = npp1; i < mrows; ++i) sm += A[i * lda + jj] * ucol[i];
                    sm += cp * up;
                    if (sm != 0.0) {
                        sm *= bb;
                        A[(nsetp - 1) * lda + jj] = cp + sm * up;
                        for (int i = npp1; i < mrows; ++i) A[i * lda + jj] += sm * ucol[i];
                    }
                }
            }
        }
        if (cx.tid == 0) w[j] = 0.0;
        for (int i = npp1 + cx.tid; i < mrows; i += cx.nthr) A[i * lda + j] = 0.0;
        cx.sync();
        // ---- loop B
        while (true) {
            int jj = -1;
            for (int ip = nsetp - 1; ip >= 0; --ip) {
                if (ip != nsetp - 1) {
                    const double zn = z[ip + 1];
                    for (int t = cx.tid; t <= ip; t += cx.nthr) z[t] -= zn * A[t * lda + jj];
                    cx.sync();
                }
                jj = idx[ip];
                const double zi = z[ip] / A[ip * lda + jj];
                cx.sync();
                if (cx.tid == 0) z[ip] = zi;
                cx.sync();
            }
            iter += 1;
            if (iter > itmax) { mode = 3; break; }
            // ---- step length: the smallest t <= 1 (the LAST position on ties, as the sequential loop does)
            double alpha = 2.0;
            int jdel = -1;
            for (int ip = cx.tid; ip < nsetp; ip += cx.nthr) {
                const double zi = z[ip];
                if (zi > 0.0) continue;
                const double xl_ = x[idx[ip]];
                const double t = -xl_ / (zi - xl_);
                if (t <= 1.0 && !(alpha < t)) { alpha = t; jdel = ip; }
            }
            {
                double key = -alpha;
                cx.argbest(key, jdel, false);                // the largest -t, LAST position on ties
                alpha = -key;
            }
            if (jdel < 0) alpha = 1.0;
            cx.sync();
            for (int ip = cx.tid; ip < nsetp; ip += cx.nthr) {
                const int l = idx[ip];
                x[l] = (1.0 - alpha) * x[l] + alpha * z[ip];
            }
            cx.sync();
            if (jdel < 0) break;
            // ---- step 11: delete column(s)
            int i = idx[jdel];
            while (true) {
                cx.sync();
                if (cx.tid == 0) x[i] = 0.0;
                for (int jc = jdel + 1; jc < nsetp; ++jc) {
                    cx.sync();
                    const int ii = idx[jc];
                    const double a = A[(jc - 1) * lda + ii], bb = A[jc * lda + ii];
                    double c, s, sig;                      // BLAS drotg
                    const double roe = fabs(a) > fabs(bb) ? a : bb;
                    const double scale = fabs(a) + fabs(bb);
                    if (scale == 0.0) { c = 1.0; s = 0.0; sig = 0.0; }
                    else {
                        sig = scale * sqrt((a / scale) * (a / scale) + (bb / scale) * (bb / scale));
                        sig = roe < 0.0 ? -sig : sig;
                        c = a / sig; s = bb / sig;
                    }
                    const double b1 = b[jc - 1], b2 = b[jc];
                    cx.sync();
                    for (int col = cx.tid; col < ncol; col += cx.nthr) {
                        const double r1 = A[(jc - 1) * lda + col], r2 = A[jc * lda + col];
                        A[(jc - 1) * lda + col] = c * r1 + s * r2;
                        A[jc * lda + col] = -s * r1 + c * r2;
                    }
                    cx.sync();
                    if (cx.tid == 0) {
                        idx[jc - 1] = ii;
                        A[(jc - 1) * lda + ii] = sig;
                        A[jc * lda + ii] = 0.0;
                        b[jc - 1] = c * b1 + s * b2;
                        b[jc] = -s * b1 + c * b2;
                    }
                }
                npp1 = nsetp - 1;
                nsetp -= 1;
                iz1 -= 1;
                cx.sync();
                if (cx.tid == 0) idx[iz1] = i;
                cx.sync();
                if (nsetp <= 0) { mode = 3; break; }
                double key = -1.0;
                int jq = -1;
                for (int q = cx.tid; q < nsetp; q += cx.nthr)
                    if (x[idx[q]] <= 0.0 && jq < 0) { jq = q; key = 1.0; }
                cx.argbest(key, jq, true);                   // the first such position
                if (!(key > 0.0) || jq < 0) break;
                jdel = jq;
                i = idx[jq];
            }
            if (mode != 1) break;
            for (int t = cx.tid; t < mrows; t += cx.nthr) z[t] = b[t];
            cx.sync();
        }
        if (mode != 1) break;
    }
    cx.sync();
    double rn = 0.0;
    if (nsetp < mrows) {
        const int k = npp1 < mrows - 1 ? npp1 : mrows - 1;
        for (int i = k + cx.tid; i < mrows; i += cx.nthr) rn += b[i] * b[i];
    }
    rnorm = sqrt(cx.sum(rn));
    if (npp1 > mrows - 1) {
        for (int jx = cx.tid; jx < ncol; jx += cx.nthr) w[jx] = 0.0;
        cx.sync();
    }
    return mode;
}

// ---------------------------------------------------------------- LSQ = LSEI -> LSI -> LDP
// Work arrays of one QP (views into the block's scratch)
struct OgsQp {
    int nq, mc, mg, ld;                // variables, equality rows, inequality rows (with bounds), leading dim
    double *E, *f, *C, *d, *ups, *binv, *G, *h, *x, *w, *E2T, *GT, *An, *bn, *zn, *xn, *wn, *ucol, *f2, *fres;
    int* idx;
    size_t ldg;                        // leading dimension of GT / An ( = mg of the shape)
    double* clk;                       // the instance's scalars (cycle counters)
};

// min ||E x - f||  s.t.  C x = d,  G x >= h.   Returns SLSQP's mode; x (nq) and w (mc + mg).
template <class Cx>
OGS_FN int ogs_lsei(Cx& cx, OgsQp& Q) {
    const int nq = Q.nq, mc = Q.mc, mg = Q.mg, ld = Q.ld, me = Q.nq;
    if (mc >= nq || mg <= 0) return 2;
    const int l = nq - mc;
    long long t0c = cx.clock();
    // ---- triangularise C from the right (C K = [C1 0]) and apply K to E and G: every row goes through the same
    //      reflections in the same order as in Lawson & Hanson's loop.
    //      Phase 1: step i defines H_i from row i of C (one warp) and applies it to the rows below (all warps,
    //      four rows per warp in flight).
    for (int i = 0; i < mc; ++i) {
        double* row = Q.C + (size_t)i * ld;
        if (cx.warp == 0) {
            double up = 0.0, binv = 0.0;
            if (i + 1 < nq) {
                double cl = 0.0;
                for (int k = i + 1 + cx.lane; k < nq; k += cx.wsize) cl = fmax(cl, fabs(row[k]));
                cl = cx.wmax(cl);
                const double upiv = row[i];
                cl = fmax(cl, fabs(upiv));
                if (cl > 0.0) {
                    const double clinv = 1.0 / cl;
                    double sm = 0.0;
                    for (int k = i + 1 + cx.lane; k < nq; k += cx.wsize) { const double t = row[k] * clinv; sm += t * t; }
                    sm = cx.wsum(sm);
                    sm += (upiv * clinv) * (upiv * clinv);
                    cl = cl * sqrt(sm);
                    if (upiv > 0.0) cl = -cl;
                    up = upiv - cl;
                    const double b = up * cl;
                    if (b < 0.0) binv = 1.0 / b;
                    cx.wsync();
                    if (cx.lane == 0) row[i] = cl;
                }
            }
            if (cx.lane == 0) { Q.ups[i] = up; Q.binv[i] = binv; }
        }
        cx.sync();
        ogs_reflect_matrix(cx, Q.C + (size_t)(i + 1) * ld, ld, mc - i - 1, row, i, nq, Q.ups[i], Q.binv[i]);
        cx.sync();
    }
    if (cx.tid == 0) Q.clk[17] += (double)(cx.clock() - t0c);        // (phase 1 alone)
    //      Phase 2: the rows of E and G are independent of each other: a warp keeps `wrows` of them in shared
    //      memory and takes them through H_0 .. H_{mc-1} without a block barrier or a global round trip per step.
    {
        const int R = cx.wrows;
        const size_t stride = cx.wstride;
        double* rb = cx.wbuf();
        for (int r0 = cx.warp * R; r0 < me + mg; r0 += cx.nwarps * R) {
            const int cnt = (me + mg - r0) < R ? (me + mg - r0) : R;
            for (int r = 0; r < R; ++r) {
                const int rr = r0 + r;
                const double* src = rr < me ? Q.E + (size_t)rr * ld : Q.G + (size_t)(rr - me) * ld;
                for (int k = cx.lane; k < nq; k += cx.wsize) rb[r * stride + k] = (r < cnt) ? src[k] : 0.0;
            }
            cx.wsync();
            if (Cx::kLanes32 && (cx.flags & 1) && R == 4 && nq <= 256) ogs_reflect_buffer_reg<4, 8>(cx, rb, stride, Q.C, ld, Q.ups, Q.binv, 0, mc, nq);
            else if (R == 4) ogs_reflect_buffer<4>(cx, rb, stride, Q.C, ld, Q.ups, Q.binv, 0, mc, nq);
            else if (R == 2) ogs_reflect_buffer<2>(cx, rb, stride, Q.C, ld, Q.ups, Q.binv, 0, mc, nq);
            else ogs_reflect_buffer<1>(cx, rb, stride, Q.C, ld, Q.ups, Q.binv, 0, mc, nq);
            for (int r = 0; r < cnt; ++r) {
                const int rr = r0 + r;
                double* dst = rr < me ? Q.E + (size_t)rr * ld : Q.G + (size_t)(rr - me) * ld;
                for (int k = cx.lane; k < nq; k += cx.wsize) dst[k] = rb[r * stride + k];
            }
            cx.wsync();
        }
    }
    cx.sync();
    if (cx.tid == 0) Q.clk[OGS_CLK_LSEI] += (double)(cx.clock() - t0c);
    t0c = cx.clock();
    // ---- solve C1 x1 = d (forward, axpy form on d)
    for (int i = 0; i < mc; ++i) {
        const double cii = Q.C[(size_t)i * ld + i];
        if (fabs(cii) < OGS_EPS) return 6;
        double sm = 0.0;
        for (int k = cx.tid; k < i; k += cx.nthr) sm += Q.C[(size_t)i * ld + k] * Q.x[k];
        sm = cx.sum(sm);
        if (cx.tid == 0) Q.x[i] = (Q.d[i] - sm) / cii;
        cx.sync();
    }
    for (int k = cx.tid; k < mc + mg; k += cx.nthr) Q.w[k] = 0.0;
    // ---- f <- f - E1 x1,  h <- h - G1 x1   (the full rotated E / G stay for the multipliers)
    double* f2 = Q.f2;
    for (int r = cx.warp; r < me; r += cx.nwarps) {
        double sm = 0.0;
        for (int k = cx.lane; k < mc; k += cx.wsize) sm += Q.E[(size_t)r * ld + k] * Q.x[k];
        sm = cx.wsum(sm);
        if (cx.lane == 0) f2[r] = Q.f[r] - sm;
    }
    for (int r = cx.warp; r < mg; r += cx.nwarps) {
        double sm = 0.0;
        for (int k = cx.lane; k < mc; k += cx.wsize) sm += Q.G[(size_t)r * ld + k] * Q.x[k];
        sm = cx.wsum(sm);
        if (cx.lane == 0) Q.GT[(size_t)l * Q.ldg + r] = Q.h[r] - sm;         // row l of GT = h2
    }
    // E2T[j][i] = E[i][mc + j],  GT[j][k] = G[k][mc + j]
    for (size_t e = cx.tid; e < (size_t)l * me; e += cx.nthr) {
        const int j = (int)(e / me), i = (int)(e % me);
        Q.E2T[(size_t)j * ld + i] = Q.E[(size_t)i * ld + mc + j];
    }
    for (size_t e = cx.tid; e < (size_t)l * mg; e += cx.nthr) {
        const int j = (int)(e / mg), k = (int)(e % mg);
        Q.GT[(size_t)j * Q.ldg + k] = Q.G[(size_t)k * ld + mc + j];
    }
    cx.sync();
    // ---- LSI: QR of E2 (columns = rows of E2T), applied to f2
    for (int i = 0; i < l; ++i) {
        bool ident;
        double* col = Q.E2T + (size_t)i * ld;
        const double up = ogs_h12_construct(cx, col, 1, i, i + 1, me, ident);
        ogs_h12_rows(cx, col, i, i + 1, me, up, ident, Q.E2T + (size_t)(i + 1) * ld, ld, l - i - 1);
        ogs_h12_vec(cx, col, 1, i, i + 1, me, up, ident, f2);
    }
    cx.sync();
    for (int j = 0; j < l; ++j)
        if (!(fabs(Q.E2T[(size_t)j * ld + j]) >= OGS_EPS)) return 5;
    // ---- G2 <- G2 R^-1 (one thread per constraint), h2 <- h2 - G2 f2
    for (int k = cx.tid; k < mg; k += cx.nthr) {
        for (int j = 0; j < l; ++j) {
            double sm = Q.GT[(size_t)j * Q.ldg + k];
            const double* rj = Q.E2T + (size_t)j * ld;            // R[t][j] = E2[t][j] = E2T[j][t]
            for (int t = 0; t < j; ++t) sm -= Q.GT[(size_t)t * Q.ldg + k] * rj[t];
            Q.GT[(size_t)j * Q.ldg + k] = sm / rj[j];
        }
        double hk = Q.GT[(size_t)l * Q.ldg + k];
        for (int j = 0; j < l; ++j) hk -= Q.GT[(size_t)j * Q.ldg + k] * f2[j];
        Q.GT[(size_t)l * Q.ldg + k] = hk;
    }
    cx.sync();
    // ---- LDP: min ||z|| s.t. G2 z >= h2, as NNLS on [G2'; h2'] u = e_{l+1}
    const int mrows = l + 1;
    for (size_t e = cx.tid; e < (size_t)mrows * mg; e += cx.nthr) {
        const int i = (int)(e / mg), k = (int)(e % mg);
        Q.An[(size_t)i * Q.ldg + k] = Q.GT[(size_t)i * Q.ldg + k];
    }
    for (int i = cx.tid; i < mrows; i += cx.nthr) Q.bn[i] = (i == l) ? 1.0 : 0.0;
    cx.sync();
    double rnorm;
    if (cx.tid == 0) Q.clk[OGS_CLK_LSI] += (double)(cx.clock() - t0c);
    t0c = cx.clock();
    int mode = ogs_nnls(cx, Q.An, Q.ldg, mrows, mg, Q.bn, Q.xn, Q.wn, Q.zn, Q.ucol, Q.idx, rnorm);
    if (cx.tid == 0) Q.clk[OGS_CLK_NNLS] += (double)(cx.clock() - t0c);
    if (mode != 1) return mode;
    if (rnorm <= 0.0) return 4;
    double hu = 0.0;
    for (int k = cx.tid; k < mg; k += cx.nthr) hu += Q.GT[(size_t)l * Q.ldg + k] * Q.xn[k];
    double fac = 1.0 - cx.sum(hu);
    if (!((1.0 + fac) - 1.0 > 0.0)) return 4;
    fac = 1.0 / fac;
    // z = fac G2' u;  x2 = R^-1 (z + f2);  multipliers w[mc + k] = fac u_k
    double* x2 = Q.x + mc;
    for (int j = cx.warp; j < l; j += cx.nwarps) {
        double sm = 0.0;
        for (int k = cx.lane; k < mg; k += cx.wsize) sm += Q.GT[(size_t)j * Q.ldg + k] * Q.xn[k];
        sm = cx.wsum(sm);
        if (cx.lane == 0) x2[j] = fac * sm + f2[j];
    }
    for (int k = cx.tid; k < mg; k += cx.nthr) Q.w[mc + k] = fac * Q.xn[k];
    cx.sync();
    for (int i = l - 1; i >= 0; --i) {                      // back substitution, axpy form: R[t][i] = E2T[i][t]
        const double xi = x2[i] / Q.E2T[(size_t)i * ld + i];
        cx.sync();
        if (cx.tid == 0) x2[i] = xi;
        for (int t = cx.tid; t < i; t += cx.nthr) x2[t] -= Q.E2T[(size_t)i * ld + t] * xi;
        cx.sync();
    }
    return 1;
}

// the part of LSEI after LSI: multipliers of the equality rows and the solution in the original variables
// (E, G are still the rotated matrices).
template <class Cx>
OGS_FN void ogs_lsei_finish(Cx& cx, OgsQp& Q) {
    double* fres = Q.fres;
    const int nq = Q.nq, mc = Q.mc, mg = Q.mg, ld = Q.ld, me = Q.nq;
    // fres = E x - f  (rotated E, rotated x)
    for (int r = cx.warp; r < me; r += cx.nwarps) {
        double sm = 0.0;
        for (int k = cx.lane; k < nq; k += cx.wsize) sm += Q.E[(size_t)r * ld + k] * Q.x[k];
        sm = cx.wsum(sm);
        if (cx.lane == 0) fres[r] = sm - Q.f[r];
    }
    cx.sync();
    // d_i = E[:, i]' fres - G[:, i]' w_ineq
    for (int i = cx.tid; i < mc; i += cx.nthr) {
        double sm = 0.0;
        for (int r = 0; r < me; ++r) sm += Q.E[(size_t)r * ld + i] * fres[r];
        for (int k = 0; k < mg; ++k) sm -= Q.G[(size_t)k * ld + i] * Q.w[mc + k];
        Q.d[i] = sm;
    }
    cx.sync();
    for (int i = mc - 1; i >= 0; --i) {
        const double up = Q.ups[i];
        ogs_h12_vec(cx, Q.C + (size_t)i * ld, 1, i, i + 1, nq, up, up == 0.0, Q.x);
    }
    for (int i = mc - 1; i >= 0; --i) {                     // w_i = (d_i - sum_{j > i} C[j][i] w_j) / C[i][i]
        double sm = 0.0;
        for (int j = i + 1 + cx.tid; j < mc; j += cx.nthr) sm += Q.C[(size_t)j * ld + i] * Q.w[j];
        sm = cx.sum(sm);
        if (cx.tid == 0) Q.w[i] = (Q.d[i] - sm) / Q.C[(size_t)i * ld + i];
        cx.sync();
    }
}

// ---------------------------------------------------------------- one instance
struct OgsInst {
    const OgsShape* S;
    double* x;                 // [n]   decision vector (in / out)
    const double* c;           // [M]   constraint values and cost at x (evaluator output)
    const double* vals;        // [nnz] packed Jacobian at x (evaluator output)
    double* st;                // persistent state
    double* W;                 // the block's scratch
};

// cost gradient (row m of the Jacobian) into g[0 .. n)
template <class Cx>
OGS_FN void ogs_cost_gradient(Cx& cx, const OgsInst& I, double* g) {
    const OgsShape& S = *I.S;
    for (int j = cx.tid; j < S.n; j += cx.nthr) g[j] = 0.0;
    cx.sync();
    for (int e = S.rowptr[S.m] + cx.tid; e < S.rowptr[S.m + 1]; e += cx.nthr) g[S.rcol[e]] = I.vals[S.rpos[e]];
    cx.sync();
}

// out[j] = sum_rows A[r][j] * r[r]   (constraint rows only), one thread per variable
template <class Cx>
OGS_FN void ogs_at_times(Cx& cx, const OgsInst& I, const double* r, double* out) {
    const OgsShape& S = *I.S;
    for (int j = cx.tid; j < S.n; j += cx.nthr) {
        double sm = 0.0;
        for (int e = S.colptr[j]; e < S.colptr[j + 1]; ++e) {
            const int row = S.prow[e];
            if (row < S.m) sm += I.vals[e] * r[row];
        }
        out[j] = sm;
    }
    cx.sync();
}

// SLSQP's LSQ on the current B = L D L', g, A, c and bounds; aug: the relaxed problem with the slack variable.
// Writes s (nq) and the multipliers r (m); returns the mode.
template <class Cx>
OGS_FN int ogs_lsq(Cx& cx, const OgsInst& I, bool aug, double rho) {
    const OgsShape& S = *I.S;
    const int n = S.n, m = S.m, meq = S.meq, mineq = S.mineq, nq = aug ? n + 1 : n, ld = S.n1;
    double* st = I.st;
    double* W = I.W;
    const double* LT = st + S.o_lt;
    const double* DG = st + S.o_dg;
    const double* g = st + S.o_g;
    OgsQp Q;
    Q.nq = nq; Q.mc = meq; Q.ld = ld; Q.ldg = S.mg;
    Q.mg = mineq + S.nlo + S.nhi + (aug ? 2 : 0);
    Q.E = W + S.o_E; Q.f = W + S.o_f; Q.C = W + S.o_C; Q.d = W + S.o_d; Q.ups = W + S.o_ups; Q.G = W + S.o_G;
    Q.h = W + S.o_h; Q.x = W + S.o_xq; Q.w = W + S.o_wq; Q.E2T = W + S.o_E2T; Q.GT = W + S.o_GT; Q.An = W + S.o_An;
    Q.bn = W + S.o_bn; Q.zn = W + S.o_zn; Q.xn = W + S.o_xn; Q.wn = W + S.o_wn; Q.ucol = W + S.o_ucol;
    Q.f2 = W + S.o_f2; Q.fres = W + S.o_fres; Q.binv = W + S.o_binv; Q.idx = (int*)(W + S.o_idx);
    Q.clk = st + S.o_sc;
    long long t0c = cx.clock();
    // ---- E = D^1/2 L' (upper triangular), f = -E^-T g
    for (size_t e = cx.tid; e < (size_t)nq * ld; e += cx.nthr) Q.E[e] = 0.0;
    for (size_t e = cx.tid; e < (size_t)meq * ld; e += cx.nthr) Q.C[e] = 0.0;
    for (size_t e = cx.tid; e < (size_t)Q.mg * ld; e += cx.nthr) Q.G[e] = 0.0;
    cx.sync();
    for (int i = cx.warp; i < n; i += cx.nwarps) {
        const double dg = sqrt(DG[i]);
        for (int j = i + cx.lane; j < n; j += cx.wsize) Q.E[(size_t)i * ld + j] = (j == i) ? dg : dg * LT[(size_t)i * n + j];
    }
    if (aug && cx.tid == 0) Q.E[(size_t)n * ld + n] = rho;
    for (int j = cx.tid; j < nq; j += cx.nthr) Q.f[j] = (j < n) ? g[j] : 0.0;
    cx.sync();
    for (int i = 0; i < nq; ++i) {                      // forward substitution with E' (axpy form)
        const double fi = Q.f[i] / Q.E[(size_t)i * ld + i];
        cx.sync();
        if (cx.tid == 0) Q.f[i] = fi;
        for (int t = i + 1 + cx.tid; t < nq; t += cx.nthr) Q.f[t] -= Q.E[(size_t)i * ld + t] * fi;
        cx.sync();
    }
    for (int j = cx.tid; j < nq; j += cx.nthr) Q.f[j] = -Q.f[j];
    // ---- C, d, G, h from the packed Jacobian (rows), the slack column, the bounds
    for (int r = cx.warp; r < m; r += cx.nwarps) {
        double* row = r < meq ? Q.C + (size_t)r * ld : Q.G + (size_t)(r - meq) * ld;
        for (int e = S.rowptr[r] + cx.lane; e < S.rowptr[r + 1]; e += cx.wsize) row[S.rcol[e]] = I.vals[S.rpos[e]];
        if (cx.lane == 0) {
            const double cr = I.c[r];
            if (r < meq) { Q.d[r] = -cr; if (aug) row[n] = -cr; }
            else { Q.h[r - meq] = -cr; if (aug) row[n] = fmax(-cr, 0.0); }
        }
    }
    {
        int k = mineq;
        for (int t = cx.tid; t < S.nlo; t += cx.nthr) {
            const int j = S.blo[t];
            Q.G[(size_t)(k + t) * ld + j] = 1.0;
            Q.h[k + t] = S.xl[j] - I.x[j];
        }
        k += S.nlo;
        if (aug) { if (cx.tid == 0) { Q.G[(size_t)k * ld + n] = 1.0; Q.h[k] = 0.0; } k += 1; }
        for (int t = cx.tid; t < S.nhi; t += cx.nthr) {
            const int j = S.bhi[t];
            Q.G[(size_t)(k + t) * ld + j] = -1.0;
            Q.h[k + t] = -(S.xu[j] - I.x[j]);
        }
        k += S.nhi;
        if (aug && cx.tid == 0) { Q.G[(size_t)k * ld + n] = -1.0; Q.h[k] = -1.0; }
    }
    cx.sync();
    if (cx.tid == 0) Q.clk[OGS_CLK_BUILD] += (double)(cx.clock() - t0c);
    int mode = ogs_lsei(cx, Q);
    if (mode != 1) return mode;
    t0c = cx.clock();
    ogs_lsei_finish(cx, Q);
    if (cx.tid == 0) Q.clk[OGS_CLK_FINISH] += (double)(cx.clock() - t0c);
    // ---- s = x clipped into the bounds, r = the multipliers of the m constraints
    double* s = st + S.o_s;
    double* r = st + S.o_r;
    for (int j = cx.tid; j < nq; j += cx.nthr) {
        double v = Q.x[j];
        if (j < n) { v = fmax(v, S.xl[j] - I.x[j]); v = fmin(v, S.xu[j] - I.x[j]); }
        else { v = fmin(fmax(v, 0.0), 1.0); }
        s[j] = v;
    }
    for (int k = cx.tid; k < m; k += cx.nthr) r[k] = Q.w[k];
    cx.sync();
    return 1;
}

// L D L' + sigma z z'  (SLSQP's ldl; LT[i][j] = L[j][i]); z is destroyed, w is scratch (n)
template <class Cx>
OGS_FN void ogs_ldl(Cx& cx, int n, double* LT, double* DG, double* z, double sigma, double* w, double* tshare) {
    if (sigma == 0.0) return;
    double t = 1.0 / sigma;
    if (sigma < 0.0) {
        for (int i = cx.tid; i < n; i += cx.nthr) w[i] = z[i];
        cx.sync();
        for (int i = 0; i < n; ++i) {
            const double v = w[i];
            t += v * v / DG[i];
            for (int j = i + 1 + cx.tid; j < n; j += cx.nthr) w[j] -= v * LT[(size_t)i * n + j];
            cx.sync();
        }
        if (t >= 0.0) t = OGS_EPS / sigma;
        if (cx.tid == 0) {
            for (int i = n - 1; i >= 0; --i) {
                const double u = w[i];
                w[i] = t;
                t -= u * u / DG[i];
            }
            *tshare = t;
        }
        cx.sync();
        t = *tshare;
    }
    for (int i = 0; i < n; ++i) {
        const double v = z[i];
        const double dgi = DG[i];
        const double delta = v / dgi;
        const double tp = sigma < 0.0 ? w[i] : t + delta * v;
        const double alpha = tp / t;
        cx.sync();
        if (cx.tid == 0) DG[i] = alpha * dgi;
        if (i == n - 1) break;
        const double beta = delta / tp;
        double* col = LT + (size_t)i * n;
        if (alpha > 4.0) {
            const double gamma = t / tp;
            for (int j = i + 1 + cx.tid; j < n; j += cx.nthr) {
                const double u = col[j];
                col[j] = gamma * u + beta * z[j];
                z[j] -= v * u;
            }
        } else {
            for (int j = i + 1 + cx.tid; j < n; j += cx.nthr) {
                z[j] -= v * col[j];
                col[j] += beta * z[j];
            }
        }
        cx.sync();
        t = tp;
    }
    cx.sync();
}

template <class Cx>
OGS_FN void ogs_reset_bfgs(Cx& cx, const OgsShape& S, double* st) {
    double* LT = st + S.o_lt;
    for (size_t e = cx.tid; e < (size_t)S.n * S.n; e += cx.nthr) LT[e] = 0.0;
    for (int i = cx.tid; i < S.n; i += cx.nthr) st[S.o_dg + i] = 1.0;
    cx.sync();
}

// sum over the constraints of max(-c_j, c_j if equality else 0), optionally weighted by mu
template <class Cx>
OGS_FN double ogs_violation(Cx& cx, const OgsShape& S, const double* c, const double* mu) {
    double sm = 0.0;
    for (int j = cx.tid; j < S.m; j += cx.nthr) {
        const double h = j < S.meq ? c[j] : 0.0;
        const double v = fmax(-c[j], h);
        sm += mu ? mu[j] * v : v;
    }
    return cx.sum(sm);
}

// One call of the reverse-communication loop for one instance.  On entry sc[MODE] says what the evaluator
// just delivered: OGS_MODE_START = c, J at x0, -1 = gradients at the accepted x, 1 = values at the trial x.
// On exit: 1 = evaluate c at x (line search), -1 = evaluate c and J at x, anything else = finished.
template <class Cx>
OGS_FN void ogs_step(Cx& cx, const OgsInst& I) {
    const OgsShape& S = *I.S;
    const int n = S.n, m = S.m, meq = S.meq;
    double* st = I.st;
    double* sc = st + S.o_sc;
    double* x0 = st + S.o_x0;
    double* s = st + S.o_s;
    double* g = st + S.o_g;
    double* mu = st + S.o_mu;
    double* r = st + S.o_r;
    double* v = st + S.o_v;
    double* u = st + S.o_u;
    double* w = st + S.o_w;
    double* LT = st + S.o_lt;
    double* DG = st + S.o_dg;
    const double acc = S.acc, tol = 10.0 * S.acc;
    const int mode_in = (int)sc[OGS_MODE];
    if (mode_in != OGS_MODE_START && mode_in != 1 && mode_in != -1) return;
    double f = I.c[m];
    double f0 = sc[OGS_F0], h3 = sc[OGS_H3], h4 = sc[OGS_H4], t0 = sc[OGS_T0], alpha = sc[OGS_ALPHA];
    int iter = (int)sc[OGS_ITER], reset = (int)sc[OGS_RESET], line = (int)sc[OGS_LINE], badlin = (int)sc[OGS_BADLIN];
    int nfev = (int)sc[OGS_NFEV], njev = (int)sc[OGS_NJEV];
    int mode_out = 0;
    bool iterate = false;                      // run (another) main iteration in this call?
    cx.sync();

    if (mode_in == OGS_MODE_START) {
        ogs_reset_bfgs(cx, S, st);
        for (int j = cx.tid; j < m; j += cx.nthr) { mu[j] = 0.0; r[j] = 0.0; }
        for (int j = cx.tid; j <= n; j += cx.nthr) s[j] = 0.0;
        ogs_cost_gradient(cx, I, g);
        iter = 0; reset = 1; nfev = 1; njev = 1; badlin = 0; f0 = f;
        iterate = true;
    } else if (mode_in == 1) {
        // ---- the values at the trial point: merit function, line search
        nfev += 1;
        double h1 = f + ogs_violation(cx, S, I.c, mu) - t0;
        bool accept = (h1 <= h3 / 10.0) || line > 10;
        if (!accept) {
            alpha = fmax(h3 / (2.0 * (h3 - h1)), OGS_ALFMIN);
            line += 1;
            h3 = alpha * h3;
            cx.sync();
            for (int j = cx.tid; j < n; j += cx.nthr) { const double sj = alpha * s[j]; s[j] = sj; I.x[j] = x0[j] + sj; }
            cx.sync();
            mode_out = 1;
        } else {
            const double viol = ogs_violation(cx, S, I.c, (const double*)nullptr);
            double ss = 0.0;
            for (int j = cx.tid; j < n; j += cx.nthr) ss += s[j] * s[j];
            ss = sqrt(cx.sum(ss));
            if ((fabs(f - f0) < acc || ss < acc) && viol < acc && !badlin && f == f) mode_out = 0;
            else mode_out = -1;
        }
    } else {
        // ---- gradients at the accepted point: damped BFGS update of L D L'
        njev += 1;
        ogs_cost_gradient(cx, I, g);
        ogs_at_times(cx, I, r, u);
        for (int j = cx.tid; j < n; j += cx.nthr) u[j] = g[j] - u[j] - v[j];
        cx.sync();
        // w = L D L' s
        double* tmp = I.W + S.o_f;                        // (scratch vectors of the QP are free here)
        for (int i = cx.warp; i < n; i += cx.nwarps) {
            double sm = 0.0;
            for (int j = i + 1 + cx.lane; j < n; j += cx.wsize) sm += LT[(size_t)i * n + j] * s[j];
            sm = cx.wsum(sm);
            if (cx.lane == 0) tmp[i] = DG[i] * (s[i] + sm);
        }
        cx.sync();
        for (int i = cx.tid; i < n; i += cx.nthr) {
            double sm = tmp[i];
            for (int j = 0; j < i; ++j) sm += LT[(size_t)j * n + i] * tmp[j];
            w[i] = sm;
        }
        cx.sync();
        double a1 = 0.0, a2 = 0.0;
        for (int j = cx.tid; j < n; j += cx.nthr) { a1 += s[j] * u[j]; a2 += s[j] * w[j]; }
        double h1 = cx.sum(a1);
        const double h2 = cx.sum(a2);
        const double h3b = 0.2 * h2;
        if (h1 < h3b) {
            const double th = (h2 - h3b) / (h2 - h1);
            h1 = h3b;
            cx.sync();
            for (int j = cx.tid; j < n; j += cx.nthr) u[j] = th * u[j] + (1.0 - th) * w[j];
            cx.sync();
        }
        if (h1 == 0.0 || h2 == 0.0) {
            reset += 1;
            if (reset > 5) {
                const double viol = ogs_violation(cx, S, I.c, (const double*)nullptr);
                double ss = 0.0;
                for (int j = cx.tid; j < n; j += cx.nthr) ss += s[j] * s[j];
                ss = sqrt(cx.sum(ss));
                mode_out = ((fabs(f - f0) < tol || ss < tol) && viol < tol && !badlin && f == f) ? 0 : 8;
            } else {
                ogs_reset_bfgs(cx, S, st);
                iterate = true;
            }
        } else {
            double* tsh = I.W + S.o_bn;
            const long long t0c = cx.clock();
            ogs_ldl(cx, n, LT, DG, u, 1.0 / h1, tmp, tsh);
            ogs_ldl(cx, n, LT, DG, w, -1.0 / h2, tmp, tsh);
            if (cx.tid == 0) sc[OGS_CLK_BFGS] += (double)(cx.clock() - t0c);
            iterate = true;
        }
    }

    while (iterate) {
        iterate = false;
        iter += 1;
        if (iter > S.itermax) { iter = S.itermax; mode_out = 9; break; }
        // ---- search direction
        h4 = 1.0;
        badlin = 0;
        int qmode = ogs_lsq(cx, I, false, 0.0);
        if (qmode == 6 && n == meq) qmode = 4;
        if (qmode == 4) {
            badlin = 1;
            double rho = 100.0;
            for (int incons = 0; incons < 6; ++incons) {
                qmode = ogs_lsq(cx, I, true, rho);
                h4 = 1.0 - s[n];
                if (qmode != 4) break;
                rho *= 10.0;
            }
        }
        if (qmode != 1) { mode_out = qmode; break; }
        // ---- multipliers, L1 test
        ogs_at_times(cx, I, r, v);
        for (int j = cx.tid; j < n; j += cx.nthr) { v[j] = g[j] - v[j]; x0[j] = I.x[j]; }
        f0 = f;
        double a1 = 0.0;
        for (int j = cx.tid; j < n; j += cx.nthr) a1 += g[j] * s[j];
        const double gs = cx.sum(a1);
        double h1 = 0.0, h2 = 0.0;
        for (int j = cx.tid; j < m; j += cx.nthr) {
            const double cj = I.c[j];
            h2 += fmax(-cj, j < meq ? cj : 0.0);
            const double ar = fabs(r[j]);
            mu[j] = fmax(ar, (mu[j] + ar) / 2.0);
            h1 += ar * fabs(cj);
        }
        h1 = fabs(gs) + cx.sum(h1);
        h2 = cx.sum(h2);
        cx.sync();
        if (h1 < acc && h2 < acc && !badlin && f == f) { mode_out = 0; break; }
        h1 = ogs_violation(cx, S, I.c, mu);
        t0 = f + h1;
        h3 = gs - h1 * h4;
        if (h3 >= 0.0) {
            reset += 1;
            if (reset > 5) {
                const double viol = ogs_violation(cx, S, I.c, (const double*)nullptr);
                double ss = 0.0;
                for (int j = cx.tid; j < n; j += cx.nthr) ss += s[j] * s[j];
                ss = sqrt(cx.sum(ss));
                mode_out = ((fabs(f - f0) < tol || ss < tol) && viol < tol && !badlin && f == f) ? 0 : 8;
                break;
            }
            ogs_reset_bfgs(cx, S, st);
            iterate = true;
            continue;
        }
        line = 1;
        alpha = 1.0;
        for (int j = cx.tid; j < n; j += cx.nthr) I.x[j] = x0[j] + s[j];
        cx.sync();
        mode_out = 1;
    }
    cx.sync();
    if (cx.tid == 0) {
        sc[OGS_F] = f; sc[OGS_F0] = f0; sc[OGS_H3] = h3; sc[OGS_H4] = h4; sc[OGS_T0] = t0; sc[OGS_ALPHA] = alpha;
        sc[OGS_MODE] = mode_out; sc[OGS_ITER] = iter; sc[OGS_RESET] = reset; sc[OGS_LINE] = line;
        sc[OGS_BADLIN] = badlin; sc[OGS_NFEV] = nfev; sc[OGS_NJEV] = njev;
    }
    cx.sync();
}
