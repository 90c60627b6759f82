// ogb_core.h -- the arithmetic of the hot path, shared by every kernel.
//
// Everything here is `__host__ __device__`: the CUDA kernels in ogb_kernels.cu call
// these functions with (threadIdx, blockDim) striding; tests/emu/ compiles the same
// header with g++ and calls them serially, so index arithmetic and row assembly can
// be checked against the golden vectors in the GPU-less build container.  The
// product library never runs this code on the host except ogb_lgl_build_host.
//
// Reference being replaced (file:line = /root/reference/OpenGoddard/optimize.py):
//   LGL basis :183-213, accessors :271-360, equality_add :670-698, cost_add :700-709,
//   Dynamics.__call__ :1122-1127, Condition :1012-1066; SciPy FD stepping
//   scipy/optimize/_numdiff.py:14-90,582-600,683-712.
#pragma once
#ifdef __CUDACC_RTC__                    // NVRTC: no host headers; math functions are built in
typedef unsigned long long uint64_t;
typedef long long int64_t;
typedef unsigned int uint32_t;
typedef int int32_t;
typedef unsigned char uint8_t;
typedef unsigned long long uintptr_t;
#else
#include <math.h>
#include <stdint.h>
#endif

#include "ogb200.h"

#if defined(__CUDACC__)
#define OGB_HD __host__ __device__ __forceinline__
#else
#define OGB_HD inline
#endif

// ------------------------------------------------------------------ device-side POD
struct OgbSec {
    int N, ns, nc, nb;      // nodes, states, controls, blocks (= ns + nc)
    int off;                // first variable of the phase in p
    int g0;                 // global node offset (phases concatenated)
    int rdef;               // first collocation-defect row of the phase in c
    int dxoff;              // offset of the phase in the D.X scratch (ns * N doubles)
    int doff;               // offset of the phase's D / D^T (N * N doubles)
    int tf_idx;             // variable index of the phase's final time
    int t0_idx;             // variable index of the phase's start time, -1 = constant t0
    int us_off;             // offset into the concatenated unit_states
    int code_off, ncode, const_off, out_off, nouts, nreg;   // node program
    int run_slot;           // output slot carrying the running-cost integrand, -1 = none
    int nnc, ncoff;         // per-node constant vectors of the node program: count, offset into P.nodec (N doubles each)
    int ng, goff;           // global variables (final times, picked states / controls) the node program reads at every
                            // node: count, offset into P.gvars
    int pad;
};
#ifdef OGB_SPEC_SECTIONS               // NVRTC build: the records are a constant table (see ogb_sec)
extern __device__ const OgbSec ogb_spec_sec[];
#endif

struct OgbKnot {            // one smooth-state knot row (optimize.py:689-696)
    int row, var_prev, var_post, pad;
    double u_prev, u_post;  // unit_states[knot][a], unit_states[knot+1][a]
};

struct OgbCol {             // how Jacobian column j is produced
    int sec;                // phase of a state/control variable; -1 = final-time variable
    int blk;                // block (state or control number) -- or the phase whose t_f it is
    int k;                  // node within the phase
    int pick;               // index into the scalar program's pick list, -1 = not picked
};

struct OgbProb {
    int nsec, n, M, meq, mineq, ndx, gtot, nknot, npick, has_running, max_nouts;
    int any_global;         // number of "global columns": decision variables some node program reads at EVERY node (a final
                            // time, a picked state); their Jacobian columns re-evaluate whole phases (0 = none)
    const double* nodec;    // per-node constant vectors of the node programs
    const int* gvars;       // global variable indices of the node programs (per phase, see OgbSec.goff)
    const int* gcvars;      // [any_global] the distinct global variables, ascending
    const int* gcol_of;     // [n] variable -> its index in gcvars, or -1
    int sc_code_off, sc_ncode, sc_const_off, sc_out_off, sc_nouts, sc_nreg, sc_cost_slot;
    double unit_time, t0x;  // t0x = time_start(0) / unit_time (optimize.py:683)
    const OgbSec* sec;
    const ogb_out* outs;    // rows already absolute
    const uint64_t* code;
    const double* consts;
    const double* D;        // row-major per phase
    const double* Dt;       // transposed per phase
    const double* w;        // LGL weights, phases concatenated
    const double* tau;      // LGL nodes, phases concatenated
    const double* ustate;
    const double* ucontrol; // unit_controls, phases concatenated (guess / trajectory kernels only)
    const OgbKnot* knots;
    const OgbCol* cols;
    const int* pickvars;
    const ogb_table* tables;    // lookup tables of OGB_INTERP
    const double* tab_x;
    const double* tab_y;
    // sparse (packed) Jacobian layout, filled in once the structure probe has run (ogb_jac_pattern):
    // entry e of an instance's packed values is J[prow[e], column of e]; column j owns entries
    // [colptr[j], colptr[j+1]) in ascending row order; pmap[j * M + r] = e, or -1 for a structural zero
    int nnz, pad_nnz;
    const int* pmap;
    const int* colptr;
    const int* prow;
};

struct OgbWork {            // per work item scratch (shared memory on the device)
    double* sp;             // [n]     decision vector (clipped when differencing)
    double* sdx;            // [ndx]   D.X at the base point
    double* sbase;          // [max_nouts][gtot] node-program outputs at the base point
    double* sc;             // [M]     c at the base point
    double* scbase;         // [sc_nouts] scalar-program outputs at the base point
    double* coef;           // [3*nsec] (tfx - tix)/2, tfx, tix per phase
    double* prefix;         // [gtot+1] left-to-right partial sums of the running cost
    double* pert;           // [max_nouts][G] node-program outputs, one perturbed column each
    double* pdx;            // [G] dx = (x0 + h) - x0
    double* px1;            // [G] x0 + h
    double* scpert;         // [sc_nouts][npick] scalar-program outputs per perturbed pick
    double* pdlt;           // [G] change of the non-dimensional state operand of D.X (0 for controls / times)
    OgbCol* pcol;           // [G] production record of each column of the work item
    double* cf;             // [ndx] (tf - t0)/2 * f at the base point, per defect row
    double* rterm;          // [gtot] running-cost terms integrand * w at the base point
    double* costp;          // [G] cost at the perturbed point of each column
    double* prdx;           // [G] 1 / dx, correctly rounded (see ogb_fd_div)
    double* gpert;          // [any_global][max_nouts][gtot] node-program outputs at every node with global variable gi
                            // perturbed (FD) / their tangents (exact), for the phases that read it
    int G;
};

// launch plan of the sweep kernel (built by ogb_host.h, passed to the kernel by value)
struct OgbPlan {
    int threads;        // CTA size of the sweep kernel (every warp produces Jacobian columns)
    int G;              // Jacobian columns per work item (perturbed-output staging capacity)
    int split;          // work items per instance = ceil(n / group); chosen per launch
    int group;          // Jacobian columns per work item (<= G)
    int TC;             // warps per CTA
    int nbuf;           // dense column buffers per warp
    unsigned long long smem_bytes;  // dynamic shared memory
    int ctas_per_sm;
    // Tail refinement (chosen per launch): instances [0, head) are cut into `split` work items of `group` columns,
    // instances [head, B) -- claimed last -- into `tsplit` items of `tgroup` columns, so the persistent CTAs finish
    // within one SMALL item of each other.  head >= B: no refinement.
    int head, tsplit, tgroup;
    // offsets (in doubles) into the dynamic shared memory block
    unsigned long long o_cache, o_sp, o_sdx, o_sbase, o_sc, o_scbase, o_coef, o_prefix, o_pert, o_pdx, o_px1,
        o_pdlt, o_pcol, o_scpert, o_cf, o_rterm, o_costp, o_prdx, o_slot, o_gpert, o_tiles, tile_stride, o_tail,
        tail_stride, o_end;
};

// Phase record s.  In the NVRTC build the records are compile-time constants
// (ogb_spec_sec, emitted by ogb_jit.h), so every field folds into an immediate.
#ifdef OGB_SPEC_SECTIONS
OGB_HD const OgbSec& ogb_sec(const OgbProb&, int s) { return ogb_spec_sec[OGB_SPEC_NSEC == 1 ? 0 : s]; }
#else
OGB_HD const OgbSec& ogb_sec(const OgbProb& P, int s) { return P.sec[s]; }
#endif

// ------------------------------------------------------------------ LGL basis
// P_n and P_n' by the three-term recurrence.
OGB_HD void ogb_legendre(int n, double x, double* P, double* dP) {
    double p0 = 1.0, p1 = x, d0 = 0.0, d1 = 1.0;
    if (n == 0) { *P = 1.0; *dP = 0.0; return; }
    for (int k = 2; k <= n; ++k) {
        double pk = ((2.0 * k - 1.0) * x * p1 - (k - 1.0) * p0) / k;
        double dk = d0 + (2.0 * k - 1.0) * p1;
        p0 = p1; p1 = pk; d0 = d1; d1 = dk;
    }
    *P = p1; *dP = d1;
}

// i-th LGL node of an N-point rule: -1, the roots of P'_{N-1}, +1 (optimize.py:183-187).
// Newton on q = P'_{N-1} with q' from Legendre's equation, started from the
// Chebyshev-Gauss-Lobatto point.
OGB_HD double ogb_lgl_node(int N, int i) {
    if (i == 0) return -1.0;
    if (i == N - 1) return 1.0;
    const int n = N - 1;
    if (2 * i == n) return 0.0;
    double x = -cos(3.14159265358979323846 * (double)i / (double)n);
    for (int it = 0; it < 100; ++it) {
        double P, dP;
        ogb_legendre(n, x, &P, &dP);
        double ddP = (2.0 * x * dP - (double)n * (n + 1.0) * P) / (1.0 - x * x);
        double step = dP / ddP;
        x -= step;
        if (fabs(step) <= 1e-16 * fabs(x)) break;
    }
    return x;
}

// w_i = 2 / (N (N-1) P_{N-1}(tau_i)^2)   (optimize.py:189-195)
OGB_HD double ogb_lgl_weight(int N, double tau_i) {
    double P, dP;
    ogb_legendre(N - 1, tau_i, &P, &dP);
    return 2.0 / ((double)N * (N - 1.0) * (P * P));
}

// D_ij (optimize.py:197-213); Pi, Pj = P_{N-1}(tau_i), P_{N-1}(tau_j)
OGB_HD double ogb_lgl_dij(int N, int i, int j, double ti, double tj, double Pi, double Pj) {
    if (i != j) return Pi / Pj / (ti - tj);
    if (i == 0) return -(double)N * (N - 1.0) * 0.25;
    if (i == N - 1) return (double)N * (N - 1.0) * 0.25;
    return 0.0;
}

// ------------------------------------------------------------------ FD step
// h for variable x0 exactly as approx_derivative(method='2-point', abs_step, bounds)
// picks it (scipy/optimize/_numdiff.py:582-600 then :46-71 with num_steps = 1).
OGB_HD double ogb_fd_step(double x0, double lb, double ub, double abs_step) {
    double h = abs_step;
    if ((x0 + h) - x0 == 0.0) {
        const double sq = 1.4901161193847656e-08;   // EPS**0.5, _eps_for_method
        h = sq * (x0 >= 0.0 ? 1.0 : -1.0) * fmax(1.0, fabs(x0));
    }
    const double lower = x0 - lb, upper = ub - x0;
    const double x1 = x0 + h;
    const bool violated = (x1 < lb) || (x1 > ub);
    const bool fitting = fabs(h) <= fmax(lower, upper);
    if (violated && fitting) h = -h;
    if (!fitting) h = (upper >= lower) ? upper : -lower;
    return h;
}

// ------------------------------------------------------------------ tape interpreter
struct OgbNodeLoad {        // node program input `a` at one node: a block of the phase, a per-node constant, a global
    const double* at;       // &sp[off + k]
    int N;
    int pblk;               // perturbed block, -1 = none
    double x1;
    int nb, nnc;            // blocks of the phase, per-node constant vectors of the program
    const double* cat;      // &nodec[ncoff + k]
    const int* gv;          // the program's global variable indices
    const double* sp;
    int pg;                 // perturbed global (index into gv), -1 = none
    double gx1;
    OGB_HD double operator()(int a) const {
        if (a < nb) return a == pblk ? x1 : at[a * N];
        a -= nb;
        if (a < nnc) return cat[a * N];
        a -= nnc;
        return a == pg ? gx1 : sp[gv[a]];
    }
};
// the loader of phase S at node k (pblk / x1: a perturbed block; pg / gx1: a perturbed global)
OGB_HD OgbNodeLoad ogb_node_load(const OgbProb& P, const double* sp, const OgbSec& S, int k, int pblk, double x1,
                                 int pg = -1, double gx1 = 0.0) {
    return OgbNodeLoad{sp + S.off + k, S.N, pblk, x1, S.nb, S.nnc, P.nodec + S.ncoff + k, P.gvars + S.goff, sp, pg, gx1};
}
// which of phase S's globals is variable j (-1: the phase does not read it)
OGB_HD int ogb_global_slot(const OgbProb& P, const OgbSec& S, int j) {
    for (int t = 0; t < S.ng; ++t)
        if (P.gvars[S.goff + t] == j) return t;
    return -1;
}

struct OgbScalarLoad {      // scalar program input: variable `a` of p
    const double* sp;
    int pvar;               // perturbed variable, -1 = none
    double x1;
    OGB_HD double operator()(int a) const { return a == pvar ? x1 : sp[a]; }
};

// scipy.interpolate.interp1d (kind='linear') at one point, reproducing SciPy's arithmetic:
// variant 0: interp1d._call_linear_np = numpy.interp (compiled_base.c arr_interp), variant 1:
// interp1d._call_linear; then interp1d._evaluate's out-of-range fill unless extrapolating.
OGB_HD double ogb_interp(const OgbProb& P, int id, double x) {
    const ogb_table T = P.tables[id];
    const double* xp = P.tab_x + T.off;
    const double* fp = P.tab_y + T.off;
    const int n = T.len;
    double y;
    if (T.variant == 0) {
        if (x > xp[n - 1]) y = fp[n - 1];
        else if (x < xp[0]) y = fp[0];
        else if (!(x == x)) y = x;
        else {
            int lo = 0, hi = n - 1;                 // last j with xp[j] <= x
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (xp[mid] <= x) lo = mid; else hi = mid - 1;
            }
            const int j = lo;
            if (j == n - 1 || xp[j] == x) y = fp[j];
            else {
                const double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
                y = slope * (x - xp[j]) + fp[j];
                if (!(y == y)) {
                    y = slope * (x - xp[j + 1]) + fp[j + 1];
                    if (!(y == y) && fp[j] == fp[j + 1]) y = fp[j];
                }
            }
        }
    } else {
        int lo = 0, hi = n;                         // searchsorted(side='left'): first i with xp[i] >= x
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (xp[mid] < x) lo = mid + 1; else hi = mid;
        }
        int i = lo < 1 ? 1 : (lo > n - 1 ? n - 1 : lo);
        const double x_lo = xp[i - 1], x_hi = xp[i];
        y = ((x - x_lo) / (x_hi - x_lo)) * fp[i] + ((x_hi - x) / (x_hi - x_lo)) * fp[i - 1];
    }
    if (!T.extrapolate) {
        if (x < xp[0]) y = T.fill_below;
        if (x > xp[n - 1]) y = T.fill_above;
    }
    return y;
}

template <class Load>
OGB_HD void ogb_run_tape(const OgbProb& P, const uint64_t* code, int ncode, const double* consts,
                         const Load& ld, double* out, int ostride) {
    double r[OGB_MAX_REG];
    for (int pc = 0; pc < ncode; ++pc) {
        const uint64_t ins = code[pc];
        const int op = (int)(ins >> 56);
        const int d = (int)((ins >> 42) & 0x3fff);
        const int a = (int)((ins >> 28) & 0x3fff);
        const int b = (int)((ins >> 14) & 0x3fff);
        const int c = (int)(ins & 0x3fff);
        double v;
        switch (op) {
            case OGB_LDP: v = ld(a); break;
            case OGB_LDC: v = consts[a]; break;
            case OGB_OUT: out[d * ostride] = r[a]; continue;
            case OGB_ADD: v = r[a] + r[b]; break;
            case OGB_SUB: v = r[a] - r[b]; break;
            case OGB_MUL: v = r[a] * r[b]; break;
            case OGB_DIV: v = r[a] / r[b]; break;
            case OGB_POW: v = pow(r[a], r[b]); break;
            case OGB_MIN: v = fmin(r[a], r[b]); break;
            case OGB_MAX: v = fmax(r[a], r[b]); break;
            case OGB_ATAN2: v = atan2(r[a], r[b]); break;
            case OGB_LT: v = r[a] < r[b] ? 1.0 : 0.0; break;
            case OGB_LE: v = r[a] <= r[b] ? 1.0 : 0.0; break;
            case OGB_GT: v = r[a] > r[b] ? 1.0 : 0.0; break;
            case OGB_GE: v = r[a] >= r[b] ? 1.0 : 0.0; break;
            case OGB_EQ: v = r[a] == r[b] ? 1.0 : 0.0; break;
            case OGB_NE: v = r[a] != r[b] ? 1.0 : 0.0; break;
            case OGB_SEL: v = r[a] != 0.0 ? r[b] : r[c]; break;
            case OGB_NEG: v = -r[a]; break;
            case OGB_SQRT: v = sqrt(r[a]); break;
            case OGB_EXP: v = exp(r[a]); break;
            case OGB_LOG: v = log(r[a]); break;
            case OGB_SIN: v = sin(r[a]); break;
            case OGB_COS: v = cos(r[a]); break;
            case OGB_TAN: v = tan(r[a]); break;
            case OGB_ABS: v = fabs(r[a]); break;
            case OGB_SQUARE: v = r[a] * r[a]; break;
            case OGB_RECIP: v = 1.0 / r[a]; break;
            case OGB_ASIN: v = asin(r[a]); break;
            case OGB_ACOS: v = acos(r[a]); break;
            case OGB_ATAN: v = atan(r[a]); break;
            case OGB_SINH: v = sinh(r[a]); break;
            case OGB_COSH: v = cosh(r[a]); break;
            case OGB_TANH: v = tanh(r[a]); break;
            case OGB_LOG10: v = log10(r[a]); break;
            case OGB_SIGN: v = (r[a] > 0.0) ? 1.0 : ((r[a] < 0.0) ? -1.0 : r[a]); break;
            case OGB_FLOOR: v = floor(r[a]); break;
            case OGB_CEIL: v = ceil(r[a]); break;
            case OGB_AND: v = (r[a] != 0.0 && r[b] != 0.0) ? 1.0 : 0.0; break;
            case OGB_OR: v = (r[a] != 0.0 || r[b] != 0.0) ? 1.0 : 0.0; break;
            case OGB_NOT: v = (r[a] == 0.0) ? 1.0 : 0.0; break;
            case OGB_INTERP: v = ogb_interp(P, b, r[a]); break;
            default: continue;
        }
        r[d] = v;
    }
}

// ------------------------------------------------------------------ exact derivatives (forward mode)
// d/dx of ogb_interp: the slope of the segment the lookup used (0 where the table clamps or fills).
OGB_HD double ogb_interp_slope(const OgbProb& P, int id, double x) {
    const ogb_table T = P.tables[id];
    const double* xp = P.tab_x + T.off;
    const double* fp = P.tab_y + T.off;
    const int n = T.len;
    if (!T.extrapolate && (x < xp[0] || x > xp[n - 1])) return 0.0;
    if (T.variant == 0) {
        if (x > xp[n - 1] || x < xp[0] || !(x == x)) return 0.0;
        int lo = 0, hi = n - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (xp[mid] <= x) lo = mid; else hi = mid - 1;
        }
        const int j = lo < n - 1 ? lo : n - 2;          // at a breakpoint: the segment to its right
        return (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
    }
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (xp[mid] < x) lo = mid + 1; else hi = mid;
    }
    const int i = lo < 1 ? 1 : (lo > n - 1 ? n - 1 : lo);
    return (fp[i] - fp[i - 1]) / (xp[i] - xp[i - 1]);
}

// One instruction of a tape in dual arithmetic: value v (already computed from a, b, c the way
// ogb_run_tape does) and tangents da, db, dc -> tangent of the result.
OGB_HD double ogb_dual_rule(const OgbProb& P, int op, int tab, double v, double a, double b, double da, double db,
                            double dc) {
    // an operation none of whose operands moves has tangent 0 whatever its value: d sqrt(s) with s = x^2 at
    // x = 0 is 0 / 0 by the chain rule, and 0 is the derivative (unary operations ignore field b)
    const bool unary = (op >= OGB_NEG && op <= OGB_CEIL) || op == OGB_NOT || op == OGB_INTERP;
    if (da == 0.0 && (unary || db == 0.0) && (op != OGB_SEL || dc == 0.0)) return 0.0;
    switch (op) {
        case OGB_ADD: return da + db;
        case OGB_SUB: return da - db;
        case OGB_MUL: return da * b + a * db;
        case OGB_DIV: return (da - v * db) / b;
        case OGB_POW: {
            double d = 0.0;
            if (da != 0.0) d = b * pow(a, b - 1.0) * da;
            if (db != 0.0) d += v * log(a) * db;
            return d;
        }
        case OGB_MIN: return (a < b || !(b == b)) ? da : db;       // fmin: a NaN operand loses
        case OGB_MAX: return (a > b || !(b == b)) ? da : db;
        case OGB_ATAN2: return (b * da - a * db) / (a * a + b * b);
        case OGB_SEL: return a != 0.0 ? db : dc;
        case OGB_NEG: return -da;
        case OGB_SQRT: return da / (2.0 * v);
        case OGB_EXP: return v * da;
        case OGB_LOG: return da / a;
        case OGB_SIN: return cos(a) * da;
        case OGB_COS: return -sin(a) * da;
        case OGB_TAN: return (1.0 + v * v) * da;
        case OGB_ABS: return a > 0.0 ? da : (a < 0.0 ? -da : 0.0);
        case OGB_SQUARE: return 2.0 * a * da;
        case OGB_RECIP: return -(v * v) * da;
        case OGB_ASIN: return da / sqrt(1.0 - a * a);
        case OGB_ACOS: return -da / sqrt(1.0 - a * a);
        case OGB_ATAN: return da / (1.0 + a * a);
        case OGB_SINH: return cosh(a) * da;
        case OGB_COSH: return sinh(a) * da;
        case OGB_TANH: return (1.0 - v * v) * da;
        case OGB_LOG10: return da / (a * 2.302585092994045684);
        case OGB_INTERP: return ogb_interp_slope(P, tab, a) * da;
        default: return 0.0;            // comparisons, logic, sign, floor, ceil: piecewise constant
    }
}

// The tape in dual arithmetic: input `seed` carries tangent 1, every other input 0; output slot t
// receives d(out_t)/d(input seed).  Same value arithmetic as ogb_run_tape.
template <class Load>
OGB_HD void ogb_run_tape_dual(const OgbProb& P, const uint64_t* code, int ncode, const double* consts,
                              const Load& ld, int seed, double* out, int ostride) {
    double r[OGB_MAX_REG], t[OGB_MAX_REG];
    for (int pc = 0; pc < ncode; ++pc) {
        const uint64_t ins = code[pc];
        const int op = (int)(ins >> 56);
        const int d = (int)((ins >> 42) & 0x3fff);
        const int a = (int)((ins >> 28) & 0x3fff);
        const int b = (int)((ins >> 14) & 0x3fff);
        const int c = (int)(ins & 0x3fff);
        double v;
        switch (op) {
            case OGB_LDP: r[d] = ld(a); t[d] = a == seed ? 1.0 : 0.0; continue;
            case OGB_LDC: r[d] = consts[a]; t[d] = 0.0; continue;
            case OGB_OUT: out[d * ostride] = t[a]; continue;
            case OGB_NOP: continue;
            case OGB_ADD: v = r[a] + r[b]; break;
            case OGB_SUB: v = r[a] - r[b]; break;
            case OGB_MUL: v = r[a] * r[b]; break;
            case OGB_DIV: v = r[a] / r[b]; break;
            case OGB_POW: v = pow(r[a], r[b]); break;
            case OGB_MIN: v = fmin(r[a], r[b]); break;
            case OGB_MAX: v = fmax(r[a], r[b]); break;
            case OGB_ATAN2: v = atan2(r[a], r[b]); break;
            case OGB_LT: v = r[a] < r[b] ? 1.0 : 0.0; break;
            case OGB_LE: v = r[a] <= r[b] ? 1.0 : 0.0; break;
            case OGB_GT: v = r[a] > r[b] ? 1.0 : 0.0; break;
            case OGB_GE: v = r[a] >= r[b] ? 1.0 : 0.0; break;
            case OGB_EQ: v = r[a] == r[b] ? 1.0 : 0.0; break;
            case OGB_NE: v = r[a] != r[b] ? 1.0 : 0.0; break;
            case OGB_SEL: v = r[a] != 0.0 ? r[b] : r[c]; break;
            case OGB_NEG: v = -r[a]; break;
            case OGB_SQRT: v = sqrt(r[a]); break;
            case OGB_EXP: v = exp(r[a]); break;
            case OGB_LOG: v = log(r[a]); break;
            case OGB_SIN: v = sin(r[a]); break;
            case OGB_COS: v = cos(r[a]); break;
            case OGB_TAN: v = tan(r[a]); break;
            case OGB_ABS: v = fabs(r[a]); break;
            case OGB_SQUARE: v = r[a] * r[a]; break;
            case OGB_RECIP: v = 1.0 / r[a]; break;
            case OGB_ASIN: v = asin(r[a]); break;
            case OGB_ACOS: v = acos(r[a]); break;
            case OGB_ATAN: v = atan(r[a]); break;
            case OGB_SINH: v = sinh(r[a]); break;
            case OGB_COSH: v = cosh(r[a]); break;
            case OGB_TANH: v = tanh(r[a]); break;
            case OGB_LOG10: v = log10(r[a]); break;
            case OGB_SIGN: v = (r[a] > 0.0) ? 1.0 : ((r[a] < 0.0) ? -1.0 : r[a]); break;
            case OGB_FLOOR: v = floor(r[a]); break;
            case OGB_CEIL: v = ceil(r[a]); break;
            case OGB_AND: v = (r[a] != 0.0 && r[b] != 0.0) ? 1.0 : 0.0; break;
            case OGB_OR: v = (r[a] != 0.0 || r[b] != 0.0) ? 1.0 : 0.0; break;
            case OGB_NOT: v = (r[a] == 0.0) ? 1.0 : 0.0; break;
            case OGB_INTERP: v = ogb_interp(P, b, r[a]); break;
            default: continue;
        }
        // (SEL passes the tangents of its value operands b, c; INTERP's table id sits in field b)
        const double tv = ogb_dual_rule(P, op, b, v, r[a], r[b], t[a], t[b], op == OGB_SEL ? t[c] : 0.0);
        r[d] = v;
        t[d] = tv;
    }
}

// ------------------------------------------------------------------ helpers
// a / dx, bit-identical to the IEEE division SciPy performs (_numdiff.py:711), from the
// correctly rounded reciprocal rdx = 1/dx shared by the whole column: q = RN(a*rdx), exact
// remainder by FMA, one correction (Markstein).  Three instructions instead of a ~25
// instruction division; differs from a true division only for non-finite / subnormal
// quotients (tools/markstein_check.py: 0 mismatches in 8e7 random cases).
OGB_HD double ogb_fd_div(double a, double dx, double rdx) {
    const double q = a * rdx;
    const double r = fma(-q, dx, a);
    return fma(r, rdx, q);
}

OGB_HD int ogb_sec_of_node(const OgbProb& P, int g) {
    int s = 0;
    while (s + 1 < P.nsec && g >= ogb_sec(P, s + 1).g0) ++s;
    return s;
}

// state_temp = states(j, i) / unit = (p * unit) / unit   (optimize.py:284 then :681)
OGB_HD double ogb_nd(double p, double u) { return (p * u) / u; }

// ------------------------------------------------------------------ K1 reference loop
// D.X for one (instance, phase, state) row; the CUDA path computes the same numbers
// with FP64 tensor-core MMAs (ogb_kernels.cu), this is the emulation / tail path.
OGB_HD void ogb_dx_row(const OgbProb& P, const OgbSec& S, int a, const double* p, double* dx) {
    const double u = P.ustate[S.us_off + a];
    const double* x = p + S.off + a * S.N;
    const double* D = P.D + S.doff;
    for (int i = 0; i < S.N; ++i) {
        double acc = 0.0;
        for (int l = 0; l < S.N; ++l) acc += D[i * S.N + l] * ogb_nd(x[l], u);
        dx[S.dxoff + a * S.N + i] = acc;
    }
}

// How the traced callbacks are executed: by the tape interpreter (ahead-of-time build), or --
// when this header is compiled by NVRTC with OGB_JIT -- by straight-line device functions
// generated from the same tapes (ogb_jit_node / ogb_jit_scalar are emitted by ogb_kernels.cu).
#ifdef OGB_JIT
template <class Load>
__device__ __forceinline__ void ogb_jit_node(const OgbProb& P, int sec, const Load& ld, double* out, int ostride);
template <class Load>
__device__ __forceinline__ void ogb_jit_scalar(const OgbProb& P, const Load& ld, double* out, int ostride);
template <class Load>
__device__ __forceinline__ void ogb_jit_node_dual(const OgbProb& P, int sec, const Load& ld, int seed, double* out, int ostride);
template <class Load>
__device__ __forceinline__ void ogb_jit_scalar_dual(const OgbProb& P, const Load& ld, int seed, double* out, int ostride);
#define OGB_NODE_PROGRAM(s, S, ld, out, stride) ogb_jit_node(P, (s), (ld), (out), (stride))
#define OGB_SCALAR_PROGRAM(ld, out, stride) ogb_jit_scalar(P, (ld), (out), (stride))
#define OGB_NODE_PROGRAM_DUAL(s, S, ld, seed, out, stride) ogb_jit_node_dual(P, (s), (ld), (seed), (out), (stride))
#define OGB_SCALAR_PROGRAM_DUAL(ld, seed, out, stride) ogb_jit_scalar_dual(P, (ld), (seed), (out), (stride))
#else
#define OGB_NODE_PROGRAM_DUAL(s, S, ld, seed, out, stride) \
    ogb_run_tape_dual(P, P.code + (S).code_off, (S).ncode, P.consts + (S).const_off, (ld), (seed), (out), (stride))
#define OGB_SCALAR_PROGRAM_DUAL(ld, seed, out, stride) \
    ogb_run_tape_dual(P, P.code + P.sc_code_off, P.sc_ncode, P.consts + P.sc_const_off, (ld), (seed), (out), (stride))
#define OGB_NODE_PROGRAM(s, S, ld, out, stride) \
    ogb_run_tape(P, P.code + (S).code_off, (S).ncode, P.consts + (S).const_off, (ld), (out), (stride))
#define OGB_SCALAR_PROGRAM(ld, out, stride) \
    ogb_run_tape(P, P.code + P.sc_code_off, P.sc_ncode, P.consts + P.sc_const_off, (ld), (out), (stride))
#endif

// ------------------------------------------------------------------ phase 2: tape jobs
// Number of tape jobs of a work item with ncols Jacobian columns (see ogb_job).
OGB_HD int ogb_njobs(const OgbProb& P, int ncols) {
    return P.gtot + 1 + ncols + (ncols > 0 ? P.any_global * P.gtot : 0);
}

// Job q of a work item: q < gtot: node program at base node q; q == gtot: scalar program
// at the base point + the per-phase time coefficients; gtot < q <= gtot + ncols: Jacobian column
// jlo + (q - gtot - 1): FD step, perturbed node program, perturbed scalar program; beyond (only when a node
// program reads a global variable): job (gi, g) = the node program at node g with global variable gi perturbed,
// for the dense column of that variable.
OGB_HD void ogb_job(const OgbProb& P, const OgbWork& W, int q, int jlo, int ncols,
                    const double* lb, const double* ub, double abs_step) {
    if (q < P.gtot) {
        const int s = ogb_sec_of_node(P, q);
        const OgbSec& S = ogb_sec(P, s);
        const OgbNodeLoad ld = ogb_node_load(P, W.sp, S, q - S.g0, -1, 0.0);
        OGB_NODE_PROGRAM(s, S, ld, W.sbase + q, P.gtot);
    } else if (q == P.gtot) {
        OgbScalarLoad ld{W.sp, -1, 0.0};
        OGB_SCALAR_PROGRAM(ld, W.scbase, 1);
        for (int s = 0; s < P.nsec; ++s) {
            const OgbSec& S = ogb_sec(P, s);
            const double tfx = ogb_nd(W.sp[S.tf_idx], P.unit_time);                 // :684
            const double tix = S.t0_idx < 0 ? P.t0x : ogb_nd(W.sp[S.t0_idx], P.unit_time);  // :683
            W.coef[3 * s + 0] = (tfx - tix) / 2.0;                                  // :686
            W.coef[3 * s + 1] = tfx;
            W.coef[3 * s + 2] = tix;
        }
    } else if (q <= P.gtot + ncols) {
        const int cl = q - P.gtot - 1;
        const int j = jlo + cl;
        const double x0 = W.sp[j];
        const double h = ogb_fd_step(x0, lb[j], ub[j], abs_step);
        const double x1 = x0 + h;                     // _numdiff.py:701
        W.px1[cl] = x1;
        W.pdx[cl] = x1 - x0;                          // _numdiff.py:707
        W.prdx[cl] = 1.0 / (x1 - x0);
        const OgbCol col = P.cols[j];
        W.pcol[cl] = col;
        double dlt = 0.0;
        if (col.sec >= 0) {
            const OgbSec& S = ogb_sec(P, col.sec);
            if (col.blk < S.ns) {
                const double u = P.ustate[S.us_off + col.blk];
                dlt = ogb_nd(x1, u) - ogb_nd(x0, u);
            }
            const OgbNodeLoad ld = ogb_node_load(P, W.sp, S, col.k, col.blk, x1);
            OGB_NODE_PROGRAM(col.sec, S, ld, W.pert + cl, W.G);
        }
        W.pdlt[cl] = dlt;
        if (col.pick >= 0) {
            OgbScalarLoad ld{W.sp, j, x1};
            OGB_SCALAR_PROGRAM(ld, W.scpert + col.pick, P.npick);
        }
    } else {
        const int e = q - (P.gtot + 1 + ncols);
        const int gi = e / P.gtot, g = e - gi * P.gtot;
        const int j = P.gcvars[gi];
        if (j < jlo || j >= jlo + ncols) return;      // that column belongs to another work item
        const int s = ogb_sec_of_node(P, g);
        const OgbSec& S = ogb_sec(P, s);
        const int slot = ogb_global_slot(P, S, j);
        if (slot < 0) return;
        const double x0 = W.sp[j];
        const double x1 = x0 + ogb_fd_step(x0, lb[j], ub[j], abs_step);
        // (a picked state / control is also a block input at its own node: perturb both views of it there)
        const OgbCol col = P.cols[j];
        const int pblk = (col.sec == s && col.k == g - S.g0) ? col.blk : -1;
        const OgbNodeLoad ld = ogb_node_load(P, W.sp, S, g - S.g0, pblk, x1, slot, x1);
        OGB_NODE_PROGRAM(s, S, ld, W.gpert + (size_t)gi * P.max_nouts * P.gtot + g, P.gtot);
    }
}

// ------------------------------------------------------------------ phase 3: c at the base point
OGB_HD void ogb_assemble_base(const OgbProb& P, const OgbWork& W, int tid, int nthr) {
    // collocation defects  D.x - (tf - t0)/2 * f   (optimize.py:686)
    for (int s = 0; s < P.nsec; ++s) {
        const OgbSec& S = ogb_sec(P, s);
        const double coef = W.coef[3 * s];
        for (int e = tid; e < S.ns * S.N; e += nthr) {
            const int a = e / S.N, i = e - a * S.N;
            const double cf = coef * W.sbase[a * P.gtot + S.g0 + i];
            W.cf[S.dxoff + e] = cf;
            W.sc[S.rdef + e] = W.sdx[S.dxoff + e] - cf;
        }
        // user rows that are pointwise in the node
        for (int slot = S.ns; slot < S.nouts; ++slot) {
            const ogb_out o = P.outs[S.out_off + slot];
            if (o.kind == OGB_OUT_RUNNING) {
                for (int k = tid; k < S.N; k += nthr)
                    W.rterm[S.g0 + k] = W.sbase[slot * P.gtot + S.g0 + k] * P.w[S.g0 + k];
                continue;
            }
            if (o.kind != OGB_OUT_EQ_POINT && o.kind != OGB_OUT_INEQ_POINT) continue;
            for (int k = tid; k < S.N; k += nthr) {
                const int g = S.g0 + k;
                if (g >= o.glo && g < o.ghi) W.sc[o.row + (g - o.glo)] = W.sbase[slot * P.gtot + g];
            }
        }
    }
    // knot rows (optimize.py:689-696; `post` is divided by the PREVIOUS phase's unit)
    for (int t = tid; t < P.nknot; t += nthr) {
        const OgbKnot K = P.knots[t];
        W.sc[K.row] = ogb_nd(W.sp[K.var_prev], K.u_prev) - (W.sp[K.var_post] * K.u_post) / K.u_prev;
    }
    // scalar user rows
    for (int slot = tid; slot < P.sc_nouts; slot += nthr) {
        const ogb_out o = P.outs[P.sc_out_off + slot];
        if (o.kind == OGB_OUT_EQ_SCALAR || o.kind == OGB_OUT_INEQ_SCALAR) W.sc[o.row] = W.scbase[slot];
    }
}

// cost = cost() + sum(running * w), summed left to right like python's sum()
// (optimize.py:700-709); one thread, after ogb_assemble_base.
OGB_HD void ogb_assemble_cost(const OgbProb& P, const OgbWork& W) {
    double cost = W.scbase[P.sc_cost_slot];
    if (P.has_running) {
        double acc = 0.0;
        for (int g = 0; g < P.gtot; ++g) {
            W.prefix[g] = acc;
            acc += W.rterm[g];
        }
        W.prefix[P.gtot] = acc;
        cost = cost + acc;
    }
    W.sc[P.M - 1] = cost;
}

// cost at the perturbed point of column cl (one thread per column): the non-integrated
// part from the perturbed scalar program if the variable is picked, the running sum redone
// left to right with the one changed term so the rounding matches the reference's sum().
// the decision variable a column record stands for, and its global-column index (-1: no node program reads it
// at every node)
OGB_HD int ogb_col_var(const OgbProb& P, const OgbCol& cd) {
    if (cd.sec < 0) return ogb_sec(P, cd.blk).tf_idx;
    const OgbSec& S = ogb_sec(P, cd.sec);
    return S.off + cd.blk * S.N + cd.k;
}
OGB_HD int ogb_col_global(const OgbProb& P, const OgbCol& cd) {
    return P.any_global ? P.gcol_of[ogb_col_var(P, cd)] : -1;
}
OGB_HD bool ogb_col_moves_cost(const OgbProb& P, const OgbCol& cd) {
    return cd.pick >= 0 || (P.has_running && (cd.sec >= 0 || ogb_col_global(P, cd) >= 0));
}
OGB_HD void ogb_cost_column(const OgbProb& P, const OgbWork& W, int cl) {
    const OgbCol cd = W.pcol[cl];
    if (!ogb_col_moves_cost(P, cd)) return;
    double cost = cd.pick >= 0 ? W.scpert[P.sc_cost_slot * P.npick + cd.pick] : W.scbase[P.sc_cost_slot];
    if (P.has_running) {
        double acc = W.prefix[P.gtot];
        const int gi = ogb_col_global(P, cd);
        if (gi >= 0) {
            // a variable the integrand may read at every node: the whole sum again, left to right, with the
            // re-evaluated terms of the phases that read it (and the one changed node of its own phase otherwise)
            const int j = ogb_col_var(P, cd);
            acc = 0.0;
            for (int s = 0; s < P.nsec; ++s) {
                const OgbSec& S = ogb_sec(P, s);
                const bool reads = ogb_global_slot(P, S, j) >= 0;
                const double* gp = W.gpert + ((size_t)gi * P.max_nouts + S.run_slot) * P.gtot;
                for (int g = S.g0; g < S.g0 + S.N; ++g) {
                    if (reads) acc += gp[g] * P.w[g];
                    else if (cd.sec == s && g == S.g0 + cd.k) acc += W.pert[S.run_slot * W.G + cl] * P.w[g];
                    else acc += W.rterm[g];
                }
            }
        } else if (cd.sec >= 0) {
            const OgbSec& S = ogb_sec(P, cd.sec);
            const int g = S.g0 + cd.k;
            acc = W.prefix[g] + W.pert[S.run_slot * W.G + cl] * P.w[g];
            for (int g2 = g + 1; g2 < P.gtot; ++g2) acc += W.rterm[g2];
        }
        cost = cost + acc;
    }
    W.costp[cl] = cost;
}

// ------------------------------------------------------------------ phase 4: one Jacobian column
// Fills the non-zeros of column j into `col` (M doubles, pre-zeroed):
//     col[r] = (c_r(x + h e_j) - c_r(x)) / dx          (_numdiff.py:708-711)
// recomputing only the rows that can change.  `lane` / `nlanes` stride the work.

// Where a column is being assembled: rows [0, meq) (user equalities, defects, knots) and
// rows [meq, M) (user inequalities, cost) may live in different buffers (the CUDA kernel
// stages them separately); tests/emu points both at one contiguous column.
struct OgbColOut {
    double* dense;          // rows [0, meq)
    double* tail;           // rows [meq, M)
    int meq;
    OGB_HD void put(int r, double v) const {
        if (r < meq) dense[r] = v; else tail[r - meq] = v;
    }
};
// ... or the packed values of the instance: entry pm[r] of vals (pm = pmap + j * M)
struct OgbColPacked {
    double* vals;
    const int* pm;
    OGB_HD void put(int r, double v) const {
        const int e = pm[r];
        if (e >= 0) vals[e] = v;
    }
};

// rows of state `a` at every node i != k: only D[i,k] * delta moves
template <class Out>
OGB_HD void ogb_scatter_drows(const OgbProb& P, const OgbWork& W, const OgbSec& S, int a, int k,
                              double dlt, double dx, double rdx, const Out& col, int lane, int nlanes) {
    const double* Dt = P.Dt + S.doff + k * S.N;                // column k of D
    for (int i = lane; i < S.N; i += nlanes) {
        if (i == k) continue;
        const int e = S.dxoff + a * S.N + i;
        const double cp = (W.sdx[e] + Dt[i] * dlt) - W.cf[e];
        col.put(S.rdef + a * S.N + i, ogb_fd_div(cp - W.sc[S.rdef + a * S.N + i], dx, rdx));
    }
}

// rows living at node k: every state's defect row (dynamics moved) and the pointwise user rows
template <class Out>
OGB_HD void ogb_scatter_noderows(const OgbProb& P, const OgbWork& W, const OgbSec& S, int sidx, int a, int k,
                                 int cl, double dlt, double dx, double rdx, const Out& col, int lane, int nlanes) {
    const double coef = W.coef[3 * sidx];
    const int g = S.g0 + k;
    for (int slot = lane; slot < S.nouts; slot += nlanes) {
        if (slot < S.ns) {
            const int e = slot * S.N + k;
            double dxp = W.sdx[S.dxoff + e];
            if (slot == a) dxp = dxp + P.D[S.doff + k * S.N + k] * dlt;
            const double cp = dxp - coef * W.pert[slot * W.G + cl];
            col.put(S.rdef + e, ogb_fd_div(cp - W.sc[S.rdef + e], dx, rdx));
        } else {
            const ogb_out o = P.outs[S.out_off + slot];
            if ((o.kind == OGB_OUT_EQ_POINT || o.kind == OGB_OUT_INEQ_POINT) && g >= o.glo && g < o.ghi) {
                const int r = o.row + (g - o.glo);
                col.put(r, ogb_fd_div(W.pert[slot * W.G + cl] - W.sc[r], dx, rdx));
            }
        }
    }
}

template <class Out>
OGB_HD void ogb_scatter_knots(const OgbProb& P, const OgbWork& W, int j, double x1, double dx,
                              double rdx, const Out& col, int lane, int nlanes) {
    for (int t = lane; t < P.nknot; t += nlanes) {
        const OgbKnot K = P.knots[t];
        if (K.var_prev == j || K.var_post == j) {
            const double xp = K.var_prev == j ? x1 : W.sp[K.var_prev];
            const double xq = K.var_post == j ? x1 : W.sp[K.var_post];
            const double cp = ogb_nd(xp, K.u_prev) - (xq * K.u_post) / K.u_prev;
            col.put(K.row, ogb_fd_div(cp - W.sc[K.row], dx, rdx));
        }
    }
}

// The phases a column moves as a whole: for a final-time column the defects of its own phase and of the next one
// rescale ((t_f - t_0) / 2 changes); for any variable a node program reads at every node (a final time in
// non-autonomous dynamics, a picked state) the reading phases are re-evaluated at every node (W.gpert).  For a
// state column of a reading phase the D-block term D[i, k] * delta is folded in here as well (the local code
// does not run for that phase).
template <class Out>
OGB_HD void ogb_scatter_phases(const OgbProb& P, const OgbWork& W, int j, const OgbCol& cd, int gi, double x1,
                               double dlt, double dx, double rdx, const Out& col, int lane, int nlanes) {
    const double tfx1 = ogb_nd(x1, P.unit_time);
    const int tsec = cd.sec < 0 ? cd.blk : -1;                   // the phase whose final time this column is
    for (int s = 0; s < P.nsec; ++s) {
        const OgbSec& S = ogb_sec(P, s);
        const bool reads = gi >= 0 && ogb_global_slot(P, S, j) >= 0;
        double coef1 = W.coef[3 * s];
        bool moved = reads;
        if (s == tsec) { coef1 = (tfx1 - W.coef[3 * s + 2]) / 2.0; moved = true; }
        else if (tsec >= 0 && S.t0_idx == j) { coef1 = (W.coef[3 * s + 1] - tfx1) / 2.0; moved = true; }
        if (!moved) continue;
        const double* f1 = reads ? W.gpert + (size_t)gi * P.max_nouts * P.gtot : W.sbase;
        const int a = (cd.sec == s && cd.blk < S.ns) ? cd.blk : -1;            // a state column of this phase
        const double* Dt = P.Dt + S.doff + (a >= 0 ? cd.k : 0) * S.N;         // column k of D
        for (int e = lane; e < S.ns * S.N; e += nlanes) {
            const int b = e / S.N, i = e - b * S.N;
            double dxp = W.sdx[S.dxoff + e];
            if (b == a) dxp = dxp + Dt[i] * dlt;
            const double cp = dxp - coef1 * f1[b * P.gtot + S.g0 + i];
            col.put(S.rdef + e, ogb_fd_div(cp - W.sc[S.rdef + e], dx, rdx));
        }
        if (!reads) continue;
        for (int slot = S.ns; slot < S.nouts; ++slot) {          // pointwise user rows of a reading phase
            const ogb_out o = P.outs[S.out_off + slot];
            if (o.kind != OGB_OUT_EQ_POINT && o.kind != OGB_OUT_INEQ_POINT) continue;
            for (int k = lane; k < S.N; k += nlanes) {
                const int g = S.g0 + k;
                if (g >= o.glo && g < o.ghi) {
                    const int r = o.row + (g - o.glo);
                    col.put(r, ogb_fd_div(f1[slot * P.gtot + g] - W.sc[r], dx, rdx));
                }
            }
        }
    }
}

// rows of the scalar program (picked variables only) and the cost row (the "+1" row)
template <class Out>
OGB_HD void ogb_scatter_scalar_cost(const OgbProb& P, const OgbWork& W, const OgbCol& cd, int cl,
                                    double dx, double rdx, const Out& col, int lane, int nlanes) {
    if (cd.pick >= 0) {
        for (int slot = lane; slot < P.sc_nouts; slot += nlanes) {
            const ogb_out o = P.outs[P.sc_out_off + slot];
            if (o.kind == OGB_OUT_EQ_SCALAR || o.kind == OGB_OUT_INEQ_SCALAR)
                col.put(o.row, ogb_fd_div(W.scpert[slot * P.npick + cd.pick] - W.sc[o.row], dx, rdx));
        }
    }
    if (lane == 0 && ogb_col_moves_cost(P, cd)) col.put(P.M - 1, ogb_fd_div(W.costp[cl] - W.sc[P.M - 1], dx, rdx));
}

template <class Out>
OGB_HD void ogb_scatter_column(const OgbProb& P, const OgbWork& W, int j, int cl,
                               const Out& col, int lane, int nlanes) {
    const OgbCol cd = W.pcol[cl];
    const double dx = W.pdx[cl], rdx = W.prdx[cl];
    const double x1 = W.px1[cl];
    const int gi = ogb_col_global(P, cd);
    const double dlt = cd.sec >= 0 ? W.pdlt[cl] : 0.0;
    if (cd.sec >= 0) {
        const OgbSec& S = ogb_sec(P, cd.sec);
        if (!(gi >= 0 && ogb_global_slot(P, S, j) >= 0)) {       // (a phase that reads j at every node: below)
            const int a = cd.blk < S.ns ? cd.blk : -1;
            if (a >= 0) ogb_scatter_drows(P, W, S, a, cd.k, dlt, dx, rdx, col, lane, nlanes);
            ogb_scatter_noderows(P, W, S, cd.sec, a, cd.k, cl, dlt, dx, rdx, col, lane, nlanes);
        }
        if (P.nknot) ogb_scatter_knots(P, W, j, x1, dx, rdx, col, lane, nlanes);
    }
    if (cd.sec < 0 || gi >= 0) ogb_scatter_phases(P, W, j, cd, gi, x1, dlt, dx, rdx, col, lane, nlanes);
    ogb_scatter_scalar_cost(P, W, cd, cl, dx, rdx, col, lane, nlanes);
}


// ================================================================== exact Jacobian
// The same rows and columns as the FD sweep, with derivatives instead of difference quotients:
//   state column (s, a, k):  defect row (s, a, i) : D[i, k]                        (i != k)
//                            defect row (s, b, k) : [b == a] D[k, k] - coef_s * d f_b / d x_a   at node k
//                            point user row at k   : d row / d x_a;   knot rows: +1 / -u_post / u_prev
//   control column          : the node rows only
//   final-time column       : defect rows of its own phase: -f / 2, of the next phase: +f / 2
//   scalar-program rows and the cost row: forward-mode tangent of the scalar tape (+ w * d running / d x)
// (reference rows: optimize.py:670-709; (p * unit) / unit has derivative 1).  The tangents come from the
// traced tapes run in dual arithmetic (ogb_run_tape_dual / the NVRTC-generated dual programs).

#define OGB_MAX_GLOBAL_OUTS 64       // output slots of a node program that reads a picked state / control (host-checked)
// Job q of a work item in exact mode: q <= gtot as in ogb_job; q > gtot: tangents of column jlo + (q - gtot - 1)
OGB_HD void ogb_job_exact(const OgbProb& P, const OgbWork& W, int q, int jlo, int ncols) {
    if (q <= P.gtot) { ogb_job(P, W, q, jlo, ncols, nullptr, nullptr, 0.0); return; }
    if (q <= P.gtot + ncols) {
        const int cl = q - P.gtot - 1;
        const int j = jlo + cl;
        const OgbCol col = P.cols[j];
        W.pcol[cl] = col;
        if (col.sec >= 0) {
            const OgbSec& S = ogb_sec(P, col.sec);
            const OgbNodeLoad ld = ogb_node_load(P, W.sp, S, col.k, -1, 0.0);
            OGB_NODE_PROGRAM_DUAL(col.sec, S, ld, col.blk, W.pert + cl, W.G);
        }
        if (col.pick >= 0) {
            OgbScalarLoad ld{W.sp, -1, 0.0};
            OGB_SCALAR_PROGRAM_DUAL(ld, j, W.scpert + col.pick, P.npick);
        }
        return;
    }
    const int e = q - (P.gtot + 1 + ncols);              // tangents with respect to a global variable, at every node
    const int gi = e / P.gtot, g = e - gi * P.gtot;
    const int j = P.gcvars[gi];
    if (j < jlo || j >= jlo + ncols) return;
    const int s = ogb_sec_of_node(P, g);
    const OgbSec& S = ogb_sec(P, s);
    const int slot = ogb_global_slot(P, S, j);
    if (slot < 0) return;
    const OgbNodeLoad ld = ogb_node_load(P, W.sp, S, g - S.g0, -1, 0.0);
    double* out = W.gpert + (size_t)gi * P.max_nouts * P.gtot + g;
    OGB_NODE_PROGRAM_DUAL(s, S, ld, S.nb + S.nnc + slot, out, P.gtot);
    // a picked state / control is also a block input at its own node: the total derivative there is the sum of
    // the tangents along both views (the column job has the block tangent in W.pert)
    const OgbCol col = P.cols[j];
    if (col.sec == s && col.k == g - S.g0) {
        double tmp[OGB_MAX_GLOBAL_OUTS];
        OGB_NODE_PROGRAM_DUAL(s, S, ld, col.blk, tmp, 1);
        for (int t = 0; t < S.nouts; ++t) out[(size_t)t * P.gtot] += tmp[t];
    }
}

// d cost / d x_j of column cl (one thread per column)
OGB_HD void ogb_cost_column_exact(const OgbProb& P, const OgbWork& W, int cl) {
    const OgbCol cd = W.pcol[cl];
    if (!ogb_col_moves_cost(P, cd)) return;
    double g = cd.pick >= 0 ? W.scpert[P.sc_cost_slot * P.npick + cd.pick] : 0.0;
    const int gi = ogb_col_global(P, cd);
    if (P.has_running && gi >= 0) {
        const int j = ogb_col_var(P, cd);
        for (int s = 0; s < P.nsec; ++s) {
            const OgbSec& S = ogb_sec(P, s);
            if (ogb_global_slot(P, S, j) >= 0) {
                const double* gp = W.gpert + ((size_t)gi * P.max_nouts + S.run_slot) * P.gtot;
                for (int q = S.g0; q < S.g0 + S.N; ++q) g = g + gp[q] * P.w[q];
            } else if (cd.sec == s) {
                g = g + W.pert[S.run_slot * W.G + cl] * P.w[S.g0 + cd.k];
            }
        }
    } else if (P.has_running && cd.sec >= 0) {
        const OgbSec& S = ogb_sec(P, cd.sec);
        g = g + W.pert[S.run_slot * W.G + cl] * P.w[S.g0 + cd.k];
    }
    W.costp[cl] = g;
}

// knot rows of column j: +1 on the previous phase's last node, -u_post / u_prev on the next phase's first
template <class Out>
OGB_HD void ogb_scatter_knots_exact(const OgbProb& P, int j, const Out& col, int lane, int nlanes) {
    for (int t = lane; t < P.nknot; t += nlanes) {
        const OgbKnot K = P.knots[t];
        if (K.var_prev == j) col.put(K.row, 1.0);
        else if (K.var_post == j) col.put(K.row, -(K.u_post / K.u_prev));
    }
}
// scalar-program rows of a picked variable and the cost row
template <class Out>
OGB_HD void ogb_scatter_scalar_cost_exact(const OgbProb& P, const OgbWork& W, const OgbCol& cd, int cl, const Out& col,
                                          int lane, int nlanes) {
    if (cd.pick >= 0) {
        for (int slot = lane; slot < P.sc_nouts; slot += nlanes) {
            const ogb_out o = P.outs[P.sc_out_off + slot];
            if (o.kind == OGB_OUT_EQ_SCALAR || o.kind == OGB_OUT_INEQ_SCALAR)
                col.put(o.row, W.scpert[slot * P.npick + cd.pick]);
        }
    }
    if (lane == 0 && ogb_col_moves_cost(P, cd)) col.put(P.M - 1, W.costp[cl]);
}

// exact counterpart of ogb_scatter_phases: d/dx_j of the rows of the phases the column moves as a whole
template <class Out>
OGB_HD void ogb_scatter_phases_exact(const OgbProb& P, const OgbWork& W, int j, const OgbCol& cd, int gi,
                                     const Out& col, int lane, int nlanes) {
    const int tsec = cd.sec < 0 ? cd.blk : -1;
    for (int s = 0; s < P.nsec; ++s) {
        const OgbSec& S = ogb_sec(P, s);
        const bool reads = gi >= 0 && ogb_global_slot(P, S, j) >= 0;
        double sign = 0.0;                               // -d coef_s / d x_j
        if (s == tsec) sign = -0.5;
        else if (tsec >= 0 && S.t0_idx == j) sign = 0.5;
        else if (!reads) continue;
        const double coef = W.coef[3 * s];
        const double* tg = W.gpert + (size_t)(gi >= 0 ? gi : 0) * P.max_nouts * P.gtot;   // tangents, if the phase reads x_j
        const int a = (cd.sec == s && cd.blk < S.ns) ? cd.blk : -1;
        const double* Dt = P.Dt + S.doff + (a >= 0 ? cd.k : 0) * S.N;
        for (int e = lane; e < S.ns * S.N; e += nlanes) {
            const int b = e / S.N, i = e - b * S.N;
            double v = sign * W.sbase[b * P.gtot + S.g0 + i];
            if (b == a) v = v + Dt[i];
            if (reads) v = v - coef * tg[b * P.gtot + S.g0 + i];
            col.put(S.rdef + e, v);
        }
        if (!reads) continue;
        for (int slot = S.ns; slot < S.nouts; ++slot) {
            const ogb_out o = P.outs[S.out_off + slot];
            if (o.kind != OGB_OUT_EQ_POINT && o.kind != OGB_OUT_INEQ_POINT) continue;
            for (int k = lane; k < S.N; k += nlanes) {
                const int g = S.g0 + k;
                if (g >= o.glo && g < o.ghi) col.put(o.row + (g - o.glo), tg[slot * P.gtot + g]);
            }
        }
    }
}

template <class Out>
OGB_HD void ogb_scatter_column_exact(const OgbProb& P, const OgbWork& W, int j, int cl, const Out& col,
                                     int lane, int nlanes) {
    const OgbCol cd = W.pcol[cl];
    const int gi = ogb_col_global(P, cd);
    if (cd.sec >= 0) {
        const OgbSec& S = ogb_sec(P, cd.sec);
        if (!(gi >= 0 && ogb_global_slot(P, S, j) >= 0)) {
            const int k = cd.k, a = cd.blk < S.ns ? cd.blk : -1;
            const double coef = W.coef[3 * cd.sec];
            if (a >= 0) {
                const double* Dt = P.Dt + S.doff + k * S.N;            // column k of D
                for (int i = lane; i < S.N; i += nlanes)
                    if (i != k) col.put(S.rdef + a * S.N + i, Dt[i]);
            }
            const int g = S.g0 + k;
            for (int slot = lane; slot < S.nouts; slot += nlanes) {
                const double t = W.pert[slot * W.G + cl];
                if (slot < S.ns) {
                    const double dkk = slot == a ? P.D[S.doff + k * S.N + k] : 0.0;
                    col.put(S.rdef + slot * S.N + k, dkk - coef * t);
                } else {
                    const ogb_out o = P.outs[S.out_off + slot];
                    if ((o.kind == OGB_OUT_EQ_POINT || o.kind == OGB_OUT_INEQ_POINT) && g >= o.glo && g < o.ghi)
                        col.put(o.row + (g - o.glo), t);
                }
            }
        }
        ogb_scatter_knots_exact(P, j, col, lane, nlanes);
    }
    if (cd.sec < 0 || gi >= 0) ogb_scatter_phases_exact(P, W, j, cd, gi, col, lane, nlanes);
    ogb_scatter_scalar_cost_exact(P, W, cd, cl, col, lane, nlanes);
}
