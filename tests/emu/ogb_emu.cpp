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
