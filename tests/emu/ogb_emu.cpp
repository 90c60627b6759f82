// ogb_emu.cpp -- TEST INFRASTRUCTURE ONLY (never linked into libogb200.so).
//
// Serial host execution of the *same* arithmetic the CUDA kernels run
// (opengoddard_b200/csrc/ogb_core.h is __host__ __device__): one "thread", one
// "lane".  It lets the GPU-less build container check tracing, tape compilation,
// layout tables and row assembly against the golden vectors before GPU time is
// spent.  The product never loads this library.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../opengoddard_b200/csrc/ogb_host.h"
#include "../../opengoddard_b200/csrc/ogb_guess.cuh"

static std::string g_err;

extern "C" {

const char* emu_last_error(void) { return g_err.c_str(); }

void* emu_problem_create(const ogb_problem_desc* d) {
    return ogb_build_host_problem(d, &g_err);
}

void emu_problem_destroy(void* h) { delete (OgbHostProblem*)h; }

int emu_problem_info_get(void* h, ogb_problem_info* o) {
    OgbHostProblem* H = (OgbHostProblem*)h;
    o->nvars = H->P.n; o->meq = H->P.meq; o->mineq = H->P.mineq; o->nrows = H->P.M;
    o->ndx = H->P.ndx; o->total_nodes = H->P.gtot; o->tile_cols = H->plan.TC;
    o->group_cols = H->plan.G; o->smem_bytes = (int)H->plan.smem_bytes;
    o->ctas_per_sm = H->plan.ctas_per_sm; o->jit = 0;
    return 0;
}

int emu_lgl_build(int N, double* tau, double* w, double* D) {
    std::vector<double> Pn(N);
    for (int i = 0; i < N; ++i) {
        double dP;
        tau[i] = ogb_lgl_node(N, i);
        ogb_legendre(N - 1, tau[i], &Pn[i], &dP);
        w[i] = ogb_lgl_weight(N, tau[i]);
    }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) D[i * N + j] = ogb_lgl_dij(N, i, j, tau[i], tau[j], Pn[i], Pn[j]);
    return 0;
}

int emu_dx_gemm(void* h, const double* p, int B, double* DX) {
    OgbHostProblem* H = (OgbHostProblem*)h;
    const OgbProb& P = H->P;
    for (int b = 0; b < B; ++b)
        for (int s = 0; s < P.nsec; ++s)
            for (int a = 0; a < P.sec[s].ns; ++a)
                ogb_dx_row(P, P.sec[s], a, p + (size_t)b * P.n, DX + (size_t)b * P.ndx);
    return 0;
}

// with_fd = 0: c only, no clipping (ogb_eval); 1: clip + c + J (ogb_eval_fd)
int emu_eval(void* h, const double* p, const double* lb, const double* ub, double abs_step,
             int B, double* c, double* J, int with_fd) {
    OgbHostProblem* H = (OgbHostProblem*)h;
    const OgbProb& P = H->P;
    const OgbPlan& pl = H->plan;
    std::vector<double> mem(pl.o_end, 0.0);
    OgbWork W;
    W.sp = mem.data() + pl.o_sp; W.sdx = mem.data() + pl.o_sdx; W.sbase = mem.data() + pl.o_sbase;
    W.sc = mem.data() + pl.o_sc; W.scbase = mem.data() + pl.o_scbase; W.coef = mem.data() + pl.o_coef;
    W.prefix = mem.data() + pl.o_prefix; W.pert = mem.data() + pl.o_pert; W.pdx = mem.data() + pl.o_pdx;
    W.px1 = mem.data() + pl.o_px1; W.scpert = mem.data() + pl.o_scpert; W.G = pl.G;
    W.pdlt = mem.data() + pl.o_pdlt; W.pcol = reinterpret_cast<OgbCol*>(mem.data() + pl.o_pcol);
    W.cf = mem.data() + pl.o_cf; W.rterm = mem.data() + pl.o_rterm; W.costp = mem.data() + pl.o_costp; W.prdx = mem.data() + pl.o_prdx;
    W.gpert = mem.data() + pl.o_gpert;
    std::vector<double> tilev(P.M);
    double* tile = tilev.data();
    std::vector<double> pclip(P.n), dxs(P.ndx);
    for (int b = 0; b < B; ++b) {
        for (int j = 0; j < P.n; ++j) {
            double x = p[(size_t)b * P.n + j];
            if (with_fd) x = fmin(fmax(x, lb[j]), ub[j]);
            pclip[j] = x;
        }
        emu_dx_gemm(h, pclip.data(), 1, dxs.data());
        const int nchunk = with_fd ? pl.split : 1;
        for (int ch = 0; ch < nchunk; ++ch) {
            const int jlo = with_fd ? ch * pl.group : 0;
            const int ncols = with_fd ? std::min(pl.group, P.n - jlo) : 0;
            for (int j = 0; j < P.n; ++j) W.sp[j] = pclip[j];
            for (int e = 0; e < P.ndx; ++e) W.sdx[e] = dxs[e];
            for (int q = 0; q < ogb_njobs(P, ncols); ++q) ogb_job(P, W, q, jlo, ncols, lb, ub, abs_step);
            ogb_assemble_base(P, W, 0, 1);
            ogb_assemble_cost(P, W);
            for (int cl = 0; cl < ncols; ++cl) ogb_cost_column(P, W, cl);
            if (ch == 0) memcpy(c + (size_t)b * P.M, W.sc, sizeof(double) * P.M);
            for (int cl = 0; cl < ncols; ++cl) {
                for (int r = 0; r < P.M; ++r) tile[r] = 0.0;
                OgbColOut out{tile, tile + P.meq, P.meq};
                ogb_scatter_column(P, W, jlo + cl, cl, out, 0, 1);
                memcpy(J + ((size_t)b * P.n + jlo + cl) * P.M, tile, sizeof(double) * P.M);
            }
        }
    }
    return 0;
}

double emu_guess_value(int kind, double t, double t0, double tf, const double* q) {
    return ogb_guess_value(kind, t, t0, tf, q);
}

void emu_philox4x32(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    ogb_philox4x32(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

double emu_u53(uint32_t a, uint32_t b) { return ogb_u53(a, b); }

// exact mode: c at clip(p) and the analytic / forward-mode Jacobian, dense J [B, n, M]
int emu_eval_exact(void* h, const double* p, const double* lb, const double* ub, int B, double* c, double* J) {
    OgbHostProblem* H = (OgbHostProblem*)h;
    const OgbProb& P = H->P;
    OgbPlan pl = H->plan;
    std::vector<double> mem(pl.o_end, 0.0);
    OgbWork W;
    W.sp = mem.data() + pl.o_sp; W.sdx = mem.data() + pl.o_sdx; W.sbase = mem.data() + pl.o_sbase;
    W.sc = mem.data() + pl.o_sc; W.scbase = mem.data() + pl.o_scbase; W.coef = mem.data() + pl.o_coef;
    W.prefix = mem.data() + pl.o_prefix; W.pert = mem.data() + pl.o_pert; W.pdx = mem.data() + pl.o_pdx;
    W.px1 = mem.data() + pl.o_px1; W.scpert = mem.data() + pl.o_scpert; W.G = pl.G;
    W.pdlt = mem.data() + pl.o_pdlt; W.pcol = reinterpret_cast<OgbCol*>(mem.data() + pl.o_pcol);
    W.cf = mem.data() + pl.o_cf; W.rterm = mem.data() + pl.o_rterm; W.costp = mem.data() + pl.o_costp; W.prdx = mem.data() + pl.o_prdx;
    W.gpert = mem.data() + pl.o_gpert;
    std::vector<double> tilev(P.M), pclip(P.n), dxs(P.ndx);
    double* tile = tilev.data();
    for (int b = 0; b < B; ++b) {
        for (int j = 0; j < P.n; ++j) pclip[j] = fmin(fmax(p[(size_t)b * P.n + j], lb[j]), ub[j]);
        emu_dx_gemm(h, pclip.data(), 1, dxs.data());
        for (int ch = 0; ch < pl.split; ++ch) {
            const int jlo = ch * pl.group;
            const int ncols = std::min(pl.group, P.n - jlo);
            for (int j = 0; j < P.n; ++j) W.sp[j] = pclip[j];
            for (int e = 0; e < P.ndx; ++e) W.sdx[e] = dxs[e];
            for (int q = 0; q < ogb_njobs(P, ncols); ++q) ogb_job_exact(P, W, q, jlo, ncols);
            ogb_assemble_base(P, W, 0, 1);
            ogb_assemble_cost(P, W);
            for (int cl = 0; cl < ncols; ++cl) ogb_cost_column_exact(P, W, cl);
            if (ch == 0) memcpy(c + (size_t)b * P.M, W.sc, sizeof(double) * P.M);
            for (int cl = 0; cl < ncols; ++cl) {
                for (int r = 0; r < P.M; ++r) tile[r] = 0.0;
                OgbColOut out{tile, tile + P.meq, P.meq};
                ogb_scatter_column_exact(P, W, jlo + cl, cl, out, 0, 1);
                memcpy(J + ((size_t)b * P.n + jlo + cl) * P.M, tile, sizeof(double) * P.M);
            }
        }
    }
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------- device SQP, serial group (csrc/ogb_sqp.h)
#include "../../opengoddard_b200/csrc/ogb_sqp_host.h"

struct EmuSqp {
    OgsHostTables T;
    int B = 0;
    std::vector<double> state, scratch, rowbuf;
};

extern "C" {

void* emu_sqp_create(int n, int m, int meq, int nnz, const int* colptr, const int* prow, const double* xl,
                     const double* xu, double acc, int itermax, int B) {
    EmuSqp* E = new EmuSqp();
    if (!ogs_build_tables(n, m, meq, nnz, colptr, prow, xl, xu, acc, itermax, E->T, &g_err)) { delete E; return nullptr; }
    OgsShape& S = E->T.S;
    S.colptr = E->T.colptr.data(); S.prow = E->T.prow.data(); S.rowptr = E->T.rowptr.data();
    S.rcol = E->T.rcol.data(); S.rpos = E->T.rpos.data(); S.blo = E->T.blo.data(); S.bhi = E->T.bhi.data();
    S.xl = E->T.xl.data(); S.xu = E->T.xu.data();
    E->B = B;
    E->state.assign((size_t)B * S.state_doubles, 0.0);
    for (int b = 0; b < B; ++b) E->state[(size_t)b * S.state_doubles + S.o_sc + OGS_MODE] = OGS_MODE_START;
    E->scratch.assign(S.scratch_doubles, 0.0);
    E->rowbuf.assign(S.n1, 0.0);
    return E;
}

void emu_sqp_destroy(void* h) { delete (EmuSqp*)h; }

// offsets (in doubles) of the persistent state: x0, s, g, mu, r, v, u, w, lt, dg, sc, total
void emu_sqp_offsets(void* h, long long* out) {
    const OgsShape& S = ((EmuSqp*)h)->T.S;
    const size_t o[] = {S.o_x0, S.o_s, S.o_g, S.o_mu, S.o_r, S.o_v, S.o_u, S.o_w, S.o_lt, S.o_dg, S.o_sc, S.state_doubles};
    for (int i = 0; i < 12; ++i) out[i] = (long long)o[i];
}

double* emu_sqp_state(void* h, int b) {
    EmuSqp* E = (EmuSqp*)h;
    return E->state.data() + (size_t)b * E->T.S.state_doubles;
}

// one reverse-communication step of every instance (x in / out)
int emu_sqp_step(void* h, double* x, const double* c, const double* vals) {
    EmuSqp* E = (EmuSqp*)h;
    const OgsShape& S = E->T.S;
    for (int b = 0; b < E->B; ++b) {
        OgsSerial cx;
        cx.wb = E->rowbuf.data(); cx.wrows = 1; cx.wstride = S.n1;
        OgsInst I{&S, x + (size_t)b * S.n, c + (size_t)b * S.M, vals + (size_t)b * S.nnz,
                  E->state.data() + (size_t)b * S.state_doubles, E->scratch.data()};
        ogs_step(cx, I);
    }
    return 0;
}

// the QP alone on the state of instance b (LT, DG, g as stored): returns the mode; s and r are in the state
int emu_sqp_lsq(void* h, int b, double* x, const double* c, const double* vals, int aug, double rho) {
    EmuSqp* E = (EmuSqp*)h;
    const OgsShape& S = E->T.S;
    OgsSerial cx;
    cx.wb = E->rowbuf.data(); cx.wrows = 1; cx.wstride = S.n1;
    OgsInst I{&S, x, c, vals, E->state.data() + (size_t)b * S.state_doubles, E->scratch.data()};
    return ogs_lsq(cx, I, aug != 0, rho);
}

}  // extern "C"
