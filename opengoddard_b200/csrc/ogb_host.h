// ogb_host.h -- host-side construction of the flat problem tables (pure C++, no CUDA).
//
// Turns an `ogb_problem_desc` (include/ogb200.h) into the arrays the kernels index:
// the decision-vector layout of the reference (`_make_param_division`,
// /root/reference/OpenGoddard/optimize.py:237-245; final times at the tail, :781), the
// row layout of `equality_add` (:670-698: user rows, then per phase / per state / per
// node defects, then knot rows), the per-column production table, the concatenated
// tapes, and the LGL basis per phase.  ogb_kernels.cu uploads them; tests/emu/ uses
// them in place.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ogb_core.h"

struct OgbHostProblem {
    std::vector<OgbSec> sec;
    std::vector<ogb_out> outs;
    std::vector<uint64_t> code;
    std::vector<double> consts, D, Dt, w, tau, ustate, ucontrol, nodec;
    std::vector<int> gvars, gcvars, gcol_of;
    std::vector<OgbKnot> knots;
    std::vector<OgbCol> cols;
    std::vector<int> pickvars;
    std::vector<ogb_table> tables;
    std::vector<double> tab_x, tab_y;
    OgbProb P;          // pointer members reference the vectors above (host view)
    OgbPlan plan;            // launch plan of the ahead-of-time (tape-interpreter) sweep kernel
    OgbPlan plan_jit;        // ... of the NVRTC build: the tapes are code there, not shared-memory data
    int jit_zero_mode = 0;   // OGB_OPT_ZERO_MODE at the time the NVRTC source is generated (compile-time there)
    std::string error;

    void bind_host() {
        P.sec = sec.data(); P.outs = outs.data(); P.code = code.data(); P.consts = consts.data();
        P.D = D.data(); P.Dt = Dt.data(); P.w = w.data(); P.ustate = ustate.data();
        P.tau = tau.data(); P.ucontrol = ucontrol.data();
        P.nodec = nodec.data(); P.gvars = gvars.data(); P.gcvars = gcvars.data(); P.gcol_of = gcol_of.data();
        P.knots = knots.data(); P.cols = cols.data(); P.pickvars = pickvars.data();
        P.tables = tables.data(); P.tab_x = tab_x.data(); P.tab_y = tab_y.data();
    }
};

static inline size_t ogb_even(size_t x) { return (x + 1) & ~(size_t)1; }

// Shared-memory plan of the sweep kernel for this problem size.
//   descriptor cache | 2 input stages (p, D.X) | base outputs | c | scalar outs | coef |
//   prefix | perturbed outputs [max_nouts][G] | dx, x1, dlt, cost [G] | column records [G] |
//   scalar perturbed outs | cf | rterm | slot table |
// resident threads per SM the sweep kernel's register budget is sized for (65536 registers / budget
// = registers per thread the NVRTC build may use).  $OGB200_THREAD_BUDGET overrides the default;
// $OGB200_JIT_MINBLOCKS (older knob) means 256 x k.
#ifndef OGB_DEFAULT_THREAD_BUDGET
#define OGB_DEFAULT_THREAD_BUDGET 768
#endif
static inline int ogb_thread_budget() {
    if (const char* tb = getenv("OGB200_THREAD_BUDGET")) {
        const int v = atoi(tb);
        if (v >= 64 && v <= 2048) return v;
    }
    const char* mb = getenv("OGB200_JIT_MINBLOCKS");
    const int k = mb ? atoi(mb) : 0;
    return (k >= 1 && k <= 8) ? 256 * k : OGB_DEFAULT_THREAD_BUDGET;
}
static inline int ogb_forced_warps() {
    const char* w = getenv("OGB200_SWEEP_WARPS");
    const int v = w ? atoi(w) : 0;
    return (v >= 1 && v <= 16) ? v : 0;
}

static inline void ogb_layout(const OgbProb& P, size_t ncode, size_t nconsts, size_t nouts, int warps,
                              OgbPlan* pl) {
    size_t o = 0;
    pl->o_cache = o;   o += ogb_even((size_t)P.nsec * sizeof(OgbSec) / 8) + ogb_even(nouts * 2) +
                            ogb_even((size_t)P.nknot * sizeof(OgbKnot) / 8) + ogb_even(ncode) +
                            ogb_even(nconsts);
    pl->o_sp = o;      o += ogb_even(P.n + 2);
    pl->o_sdx = o;     o += ogb_even(P.ndx + 2);
    o += ogb_even(P.n + 2) + ogb_even(P.ndx + 2);          // second input stage
    pl->o_sbase = o;   o += ogb_even((size_t)P.max_nouts * P.gtot);
    pl->o_sc = o;      o += ogb_even(P.M);
    pl->o_scbase = o;  o += ogb_even(P.sc_nouts + 1);
    pl->o_coef = o;    o += ogb_even(3 * P.nsec);
    pl->o_prefix = o;  o += ogb_even(P.gtot + 1);
    pl->o_pert = o;    o += ogb_even((size_t)P.max_nouts * pl->G);
    pl->o_pdx = o;     o += ogb_even(pl->G);
    pl->o_px1 = o;     o += ogb_even(pl->G);
    pl->o_pdlt = o;    o += ogb_even(pl->G);
    pl->o_pcol = o;    o += (size_t)pl->G * 2;
    pl->o_scpert = o;  o += ogb_even((size_t)P.sc_nouts * std::max(1, P.npick));
    pl->o_cf = o;      o += ogb_even(P.ndx);
    pl->o_rterm = o;   o += ogb_even(P.gtot);
    pl->o_costp = o;   o += ogb_even(pl->G);
    pl->o_prdx = o;    o += ogb_even(pl->G);
    pl->o_slot = o;    o += nouts * 2;                     // int4 per output slot
    pl->o_gpert = o;   o += ogb_even((size_t)P.any_global * P.max_nouts * P.gtot);
    pl->o_tiles = pl->o_tail = o;                          // (no column staging: J is written directly)
    pl->tile_stride = pl->tail_stride = 0;
    pl->o_end = o;
    pl->smem_bytes = o * 8 + 32;   // + 2 mbarriers + two next-item slots
    pl->TC = warps;
    pl->threads = warps * 32;
    pl->nbuf = 2;
}

static inline bool ogb_make_plan(const OgbProb& P, size_t ncode, size_t nconsts, size_t nouts,
                                 OgbPlan* pl, std::string* err, int force_warps = 0, int gmax = 0) {
    const size_t SMEM_MAX = 227 * 1024, SM_SMEM = 228 * 1024;
    const int GMAX = gmax > 0 ? gmax : 256;          // Jacobian columns staged per work item (measured: tools/sweep_group.py)
    pl->split = (P.n + GMAX - 1) / GMAX;
    pl->G = (P.n + pl->split - 1) / pl->split;
    pl->head = 0x7fffffff; pl->tsplit = pl->split; pl->tgroup = pl->G;
    pl->group = pl->G;
    // pick the CTA size (2..8 warps) that keeps the most warps resident per SM; registers
    // allow 768 threads per SM (<= 85 registers per thread)
    int best_w = 0, best_res = 0;
    if (!force_warps) force_warps = ogb_forced_warps();
    for (int w = std::max(8, force_warps); w >= 1; --w) {
        if (force_warps ? w != force_warps : (w & 1)) continue;
        ogb_layout(P, ncode, nconsts, nouts, w, pl);
        if (pl->smem_bytes > SMEM_MAX) continue;
        int ctas = (int)std::min<size_t>(16, SM_SMEM / (pl->smem_bytes + 1024));
        ctas = std::min(ctas, ogb_thread_budget() / (w * 32));
        if (ctas < 1) continue;
        const int res = ctas * w;
        if (res > best_res) { best_res = res; best_w = w; }
    }
    if (!best_w) {
        ogb_layout(P, ncode, nconsts, nouts, 2, pl);
        char b[160];
        snprintf(b, sizeof b, "problem needs %zu B of shared memory per CTA (> %zu)", pl->smem_bytes,
                 SMEM_MAX);
        *err = b;
        return false;
    }
    ogb_layout(P, ncode, nconsts, nouts, best_w, pl);
    pl->ctas_per_sm = std::max(1, std::min((int)(SM_SMEM / (pl->smem_bytes + 1024)), ogb_thread_budget() / pl->threads));
    return true;
}

static inline OgbHostProblem* ogb_build_host_problem(const ogb_problem_desc* d, std::string* err) {
    auto fail = [&](const char* m) -> OgbHostProblem* { *err = m; return nullptr; };
    if (!d || d->nsec < 1) return fail("bad descriptor: nsec < 1");
    OgbHostProblem* H = new OgbHostProblem();
    OgbProb& P = H->P;
    memset(&P, 0, sizeof P);
    P.nsec = d->nsec;
    P.unit_time = d->unit_time;
    P.t0x = d->t0 / d->unit_time;
    P.has_running = d->has_running_cost;

    // ---- decision-vector layout and LGL bases
    int off = 0, g0 = 0, dxoff = 0, doff = 0, usoff = 0;
    for (int s = 0; s < d->nsec; ++s) {
        OgbSec S;
        memset(&S, 0, sizeof S);
        S.N = d->nodes_h[s]; S.ns = d->nstates_h[s]; S.nc = d->ncontrols_h[s];
        if (S.N < 3 || S.ns < 1 || S.nc < 0) { delete H; return fail("bad descriptor: need nodes >= 3, nstates >= 1"); }
        S.nb = S.ns + S.nc;
        S.off = off; S.g0 = g0; S.dxoff = dxoff; S.doff = doff; S.us_off = usoff;
        off += S.nb * S.N; g0 += S.N; dxoff += S.ns * S.N; doff += S.N * S.N; usoff += S.ns;
        H->sec.push_back(S);
    }
    P.n = off + d->nsec;
    P.gtot = g0;
    P.ndx = dxoff;
    if (P.n > OGB_MAX_FIELD) { delete H; return fail("too many variables for the 14-bit tape operand field"); }
    H->ustate.assign(d->unit_states_h, d->unit_states_h + usoff);
    {
        int ncs = 0;
        for (const OgbSec& S : H->sec) ncs += S.nc;
        if (d->unit_controls_h) H->ucontrol.assign(d->unit_controls_h, d->unit_controls_h + ncs);
        else H->ucontrol.assign((size_t)std::max(1, ncs), 1.0);
        if (H->ucontrol.empty()) H->ucontrol.push_back(1.0);
    }
    H->D.resize(doff); H->Dt.resize(doff); H->w.resize(g0); H->tau.resize(g0);
    for (int s = 0; s < d->nsec; ++s) {
        OgbSec& S = H->sec[s];
        S.tf_idx = P.n - d->nsec + s;                       // time_final(s), optimize.py:359-360
        S.t0_idx = s == 0 ? -1 : P.n - d->nsec + s - 1;     // time_start(s), optimize.py:343-347
        const int N = S.N;
        std::vector<double> Pn(N);
        for (int i = 0; i < N; ++i) {
            double tau = ogb_lgl_node(N, i), dP;
            H->tau[S.g0 + i] = tau;
            ogb_legendre(N - 1, tau, &Pn[i], &dP);
            H->w[S.g0 + i] = ogb_lgl_weight(N, tau);
        }
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) {
                double v = ogb_lgl_dij(N, i, j, H->tau[S.g0 + i], H->tau[S.g0 + j], Pn[i], Pn[j]);
                H->D[S.doff + i * N + j] = v;
                H->Dt[S.doff + j * N + i] = v;
            }
    }

    // ---- row layout of c = [c_eq ; c_ineq ; cost]
    int row = d->meq_user;
    for (auto& S : H->sec) { S.rdef = row; row += S.ns * S.N; }
    for (int k = 0; k + 1 < d->nsec; ++k) {
        const OgbSec &A = H->sec[k], &B = H->sec[k + 1];
        if (A.ns != B.ns || !d->knot_smooth_h || !d->knot_smooth_h[k]) continue;
        for (int a = 0; a < A.ns; ++a) {
            OgbKnot K;
            memset(&K, 0, sizeof K);
            K.row = row++;
            K.var_prev = A.off + a * A.N + A.N - 1;
            K.var_post = B.off + a * B.N;
            K.u_prev = H->ustate[A.us_off + a];
            K.u_post = H->ustate[B.us_off + a];
            H->knots.push_back(K);
        }
    }
    P.nknot = (int)H->knots.size();
    P.meq = row;
    P.mineq = d->mineq_user;
    P.M = P.meq + P.mineq + 1;
    if ((double)P.n * (double)P.M >= 4.0e9) { delete H; return fail("Jacobian of one instance exceeds 2^32 entries"); }

    // ---- lookup tables
    for (int t = 0; t < d->ntables; ++t) {
        const ogb_table T = d->tables_h[t];
        if (T.len < 2 || T.off < 0) { delete H; return fail("bad lookup table"); }
        for (int k = 1; k < T.len; ++k)
            if (!(d->table_x_h[T.off + k] >= d->table_x_h[T.off + k - 1])) { delete H; return fail("lookup table abscissae must ascend"); }
        H->tables.push_back(T);
        const int end = T.off + T.len;
        if ((int)H->tab_x.size() < end) { H->tab_x.resize(end); H->tab_y.resize(end); }
        for (int k = T.off; k < end; ++k) { H->tab_x[k] = d->table_x_h[k]; H->tab_y[k] = d->table_y_h[k]; }
    }
    const int ntables = d->ntables;
    // ---- tapes
    auto add_prog = [&](const ogb_program& pr, int* code_off, int* const_off, int* out_off) -> bool {
        if (pr.nreg > OGB_MAX_REG) { *err = "tape needs more than OGB_MAX_REG registers"; return false; }
        *code_off = (int)H->code.size(); *const_off = (int)H->consts.size(); *out_off = (int)H->outs.size();
        H->code.insert(H->code.end(), pr.code_h, pr.code_h + pr.ncode);
        H->consts.insert(H->consts.end(), pr.consts_h, pr.consts_h + pr.nconsts);
        for (int i = 0; i < pr.nouts; ++i) {
            ogb_out o = pr.outs_h[i];
            switch (o.kind) {
                case OGB_OUT_EQ_POINT:
                    if (o.row < 0 || o.row + (o.ghi - o.glo) > d->meq_user) { *err = "eq point rows out of range"; return false; }
                    break;
                case OGB_OUT_EQ_SCALAR:
                    if (o.row < 0 || o.row >= d->meq_user) { *err = "eq scalar row out of range"; return false; }
                    break;
                case OGB_OUT_INEQ_POINT:
                    if (o.row < 0 || o.row + (o.ghi - o.glo) > d->mineq_user) { *err = "ineq point rows out of range"; return false; }
                    o.row += P.meq;
                    break;
                case OGB_OUT_INEQ_SCALAR:
                    if (o.row < 0 || o.row >= d->mineq_user) { *err = "ineq scalar row out of range"; return false; }
                    o.row += P.meq;
                    break;
                case OGB_OUT_COST: o.row = P.M - 1; break;
                default: break;
            }
            H->outs.push_back(o);
        }
        for (int i = 0; i < pr.ncode; ++i) {
            const uint64_t ins = pr.code_h[i];
            const int op = (int)(ins >> 56), dd = (int)((ins >> 42) & 0x3fff), aa = (int)((ins >> 28) & 0x3fff);
            if (op >= OGB_OP_COUNT) { *err = "unknown opcode in tape"; return false; }
            if (op == OGB_OUT && dd >= pr.nouts) { *err = "tape OUT slot out of range"; return false; }
            if (op == OGB_LDC && aa >= pr.nconsts) { *err = "tape constant index out of range"; return false; }
            if (op != OGB_OUT && op != OGB_NOP && dd >= pr.nreg) { *err = "tape register out of range"; return false; }
            if (op == OGB_INTERP && (int)((ins >> 14) & 0x3fff) >= ntables) { *err = "tape references a missing lookup table"; return false; }
        }
        return true;
    };
    P.max_nouts = 1;
    for (int s = 0; s < d->nsec; ++s) {
        OgbSec& S = H->sec[s];
        const ogb_program& pr = d->node_prog_h[s];
        if (!add_prog(pr, &S.code_off, &S.const_off, &S.out_off)) { delete H; return nullptr; }
        S.ncode = pr.ncode; S.nouts = pr.nouts; S.nreg = pr.nreg; S.run_slot = -1;
        if (pr.nouts < S.ns) { delete H; return fail("node program must output every state derivative"); }
        for (int i = 0; i < pr.nouts; ++i) {
            const ogb_out& o = pr.outs_h[i];
            if (i < S.ns && (o.kind != OGB_OUT_DYN || o.row != i)) { delete H; return fail("node program outputs 0..ns-1 must be the dynamics in state order"); }
            if (i >= S.ns && o.kind == OGB_OUT_DYN) { delete H; return fail("duplicate dynamics output"); }
            if (o.kind == OGB_OUT_RUNNING) S.run_slot = i;
            if (o.kind == OGB_OUT_EQ_SCALAR || o.kind == OGB_OUT_INEQ_SCALAR || o.kind == OGB_OUT_COST) { delete H; return fail("scalar output kind in a node program"); }
        }
        if (pr.n_nodec < 0 || pr.nglobals < 0 || (pr.n_nodec > 0 && !pr.nodec_h) || (pr.nglobals > 0 && !pr.globals_h)) { delete H; return fail("bad node-constant / global tables of a node program"); }
        S.nnc = pr.n_nodec; S.ncoff = (int)H->nodec.size();
        H->nodec.insert(H->nodec.end(), pr.nodec_h, pr.nodec_h + (size_t)pr.n_nodec * S.N);
        S.ng = pr.nglobals; S.goff = (int)H->gvars.size();
        for (int i = 0; i < pr.nglobals; ++i) {
            const int v = pr.globals_h[i];
            if (v < 0 || v >= P.n) { delete H; return fail("a node program's global variable is outside the decision vector"); }
            if (v < P.n - d->nsec && pr.nouts > OGB_MAX_GLOBAL_OUTS) { delete H; return fail("a node program that reads a picked state / control has too many outputs"); }
            H->gvars.push_back(v);
            H->gcvars.push_back(v);
        }
        for (int i = 0; i < pr.ncode; ++i)
            if ((int)(pr.code_h[i] >> 56) == OGB_LDP && (int)((pr.code_h[i] >> 28) & 0x3fff) >= S.nb + S.nnc + S.ng) { delete H; return fail("node program reads an input outside the phase's blocks, constants and globals"); }
        if (d->has_running_cost && S.run_slot < 0) { delete H; return fail("running cost declared but a phase has no integrand output"); }
        P.max_nouts = std::max(P.max_nouts, pr.nouts);
    }
    {
        const ogb_program& pr = *d->scalar_prog_h;
        if (!add_prog(pr, &P.sc_code_off, &P.sc_const_off, &P.sc_out_off)) { delete H; return nullptr; }
        P.sc_ncode = pr.ncode; P.sc_nouts = pr.nouts; P.sc_nreg = pr.nreg; P.sc_cost_slot = -1;
        for (int i = 0; i < pr.nouts; ++i) {
            const int k = pr.outs_h[i].kind;
            if (k == OGB_OUT_COST) P.sc_cost_slot = i;
            else if (k != OGB_OUT_EQ_SCALAR && k != OGB_OUT_INEQ_SCALAR) { delete H; return fail("node output kind in the scalar program"); }
        }
        if (P.sc_cost_slot < 0) { delete H; return fail("scalar program has no cost output"); }
        for (int i = 0; i < pr.ncode; ++i)
            if ((int)(pr.code_h[i] >> 56) == OGB_LDP) {
                const int v = (int)((pr.code_h[i] >> 28) & 0x3fff);
                if (v >= P.n) { delete H; return fail("scalar program reads a variable outside p"); }
                H->pickvars.push_back(v);
            }
        std::sort(H->pickvars.begin(), H->pickvars.end());
        H->pickvars.erase(std::unique(H->pickvars.begin(), H->pickvars.end()), H->pickvars.end());
        P.npick = (int)H->pickvars.size();
    }

    // ---- per-column production table
    H->cols.resize(P.n);
    for (int s = 0; s < d->nsec; ++s) {
        const OgbSec& S = H->sec[s];
        for (int b = 0; b < S.nb; ++b)
            for (int k = 0; k < S.N; ++k) H->cols[S.off + b * S.N + k] = OgbCol{s, b, k, -1};
        H->cols[S.tf_idx] = OgbCol{-1, s, 0, -1};
    }
    for (int i = 0; i < P.npick; ++i) H->cols[H->pickvars[i]].pick = i;
    if (H->pickvars.empty()) H->pickvars.push_back(0);   // keep the device array non-empty
    if (H->nodec.empty()) H->nodec.push_back(0.0);
    std::sort(H->gcvars.begin(), H->gcvars.end());
    H->gcvars.erase(std::unique(H->gcvars.begin(), H->gcvars.end()), H->gcvars.end());
    P.any_global = (int)H->gcvars.size();
    H->gcol_of.assign((size_t)P.n, -1);
    for (int i = 0; i < P.any_global; ++i) H->gcol_of[H->gcvars[i]] = i;
    if (H->gvars.empty()) H->gvars.push_back(0);
    if (H->gcvars.empty()) H->gcvars.push_back(0);

    H->bind_host();
    if (!ogb_make_plan(P, H->code.size(), H->consts.size(), H->outs.size(), &H->plan, err)) { delete H; return nullptr; }
    if (!ogb_make_plan(P, 0, 0, H->outs.size(), &H->plan_jit, err)) { delete H; return nullptr; }
    return H;
}
