// ogb_sweep.cuh -- K2, the fused perturbed-sweep kernel (device code only).
//
// Compiled twice from this one source:
//   * ahead of time by nvcc into libogb200.so, where the traced user callbacks are run by the
//     tape interpreter (ogb_run_tape), and
//   * at problem-creation time by NVRTC (ogb_kernels.cu: ogb_jit_*), with OGB_JIT defined and
//     the tapes lowered to straight-line device functions (ogb_jit_node / ogb_jit_scalar), so the
//     same kernel runs the user's dynamics as compiled code.  Both variants execute the same
//     floating-point operations in the same order (-fmad=false) and agree bit for bit.
#pragma once
#include "ogb_core.h"

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "OGB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra OGB_DONE_%=;\n"
        "bra OGB_WAIT_%=;\n"
        "OGB_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// TMA 1-D bulk copy shared -> global, tracked by bulk async-groups
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D = A(8x4, row) * B(4x8, col) + C, all FP64: one DMMA per warp
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// ------------------------------------------------------------------ K2: fused sweep
// Persistent CTAs; one work item = (instance, group of <= G Jacobian columns).
template <class T>
__device__ __forceinline__ const T* cache_copy(double*& cur, const T* src, size_t count, int tid, int nthr) {
    // copy `count` T's (sizeof(T) % 8 == 0) into shared memory at `cur`; returns the shared copy
    const size_t nd = (count * sizeof(T) + 7) / 8;
    const double* s64 = reinterpret_cast<const double*>(src);
    for (size_t e = tid; e < nd; e += nthr) cur[e] = s64[e];
    const T* out = reinterpret_cast<const T*>(cur);
    cur += (nd + 1) & ~(size_t)1;
    return out;
}

struct OgbSlot { int rbase, klo, khi, isdyn; };   // where output slot t of a node program lands
template <bool V> struct OgbBool { static constexpr bool value = V; };

#define OGB_FAST_MAXN 128      // register-cached row constants cover phases up to 128 nodes

// NR = ceil(max nodes per phase / 32): row constants held per lane (0 = generic column code only)
#ifndef OGB_MIN_BLOCKS
#define OGB_MIN_BLOCKS 3            // resident CTAs per SM the register budget is sized for
#endif
#ifndef OGB_MAX_THREADS
#define OGB_MAX_THREADS 256
#endif
// PACKED_OUT = 0: c only / the dense J (zero stream + overwrites) / the probes; 1: the structurally
// non-zero entries of the FD Jacobian into vals [B, nnz] (with_fd = 6); 2: the same entries of the EXACT
// Jacobian (forward-mode tangents of the tapes + the analytic D-block, with_fd = 7).  Separate kernels rather
// than run-time branches so that each gets its own register allocation (the dense column loop is
// latency-bound and spill-free).
template <int NR, int PACKED_OUT>
__global__ void __launch_bounds__(OGB_MAX_THREADS, OGB_MIN_BLOCKS)
ogb_sweep_kernel(OgbProb P, OgbPlan pl, const double* __restrict__ p, const double* __restrict__ DX,
                 const double* __restrict__ lb, const double* __restrict__ ub, double abs_step,
                 int B, double* __restrict__ c, double* __restrict__ J, int with_fd,
                 int ncode, int nconsts, int nouts, int force_generic, unsigned long long* ticket,
                 unsigned long long ticket_base, int zero_mode) {
    extern __shared__ __align__(16) double smem[];
#ifdef OGB_SPEC_M                     // NVRTC build: the problem's sizes are compile-time constants
    P.M = OGB_SPEC_M; P.n = OGB_SPEC_NVARS; P.meq = OGB_SPEC_MEQ; P.mineq = OGB_SPEC_MINEQ;
    P.gtot = OGB_SPEC_GTOT; P.ndx = OGB_SPEC_NDX; P.nsec = OGB_SPEC_NSEC; P.nknot = OGB_SPEC_NKNOT;
    P.npick = OGB_SPEC_NPICK; P.has_running = OGB_SPEC_RUNNING; P.sc_nouts = OGB_SPEC_SC_NOUTS;
    P.sc_cost_slot = OGB_SPEC_SC_COST_SLOT; P.max_nouts = OGB_SPEC_MAX_NOUTS; P.any_global = OGB_SPEC_ANY_GLOBAL;
    pl.G = OGB_SPEC_G;
#ifdef OGB_SPEC_O_SC                  // ... and so are the offsets of the shared-memory layout
    pl.o_sbase = OGB_SPEC_O_SBASE; pl.o_sc = OGB_SPEC_O_SC; pl.o_scbase = OGB_SPEC_O_SCBASE;
    pl.o_coef = OGB_SPEC_O_COEF; pl.o_prefix = OGB_SPEC_O_PREFIX; pl.o_pert = OGB_SPEC_O_PERT;
    pl.o_pdx = OGB_SPEC_O_PDX; pl.o_px1 = OGB_SPEC_O_PX1; pl.o_pdlt = OGB_SPEC_O_PDLT;
    pl.o_pcol = OGB_SPEC_O_PCOL; pl.o_scpert = OGB_SPEC_O_SCPERT; pl.o_cf = OGB_SPEC_O_CF;
    pl.o_rterm = OGB_SPEC_O_RTERM; pl.o_costp = OGB_SPEC_O_COSTP; pl.o_prdx = OGB_SPEC_O_PRDX;
    pl.o_slot = OGB_SPEC_O_SLOT; pl.o_sp = OGB_SPEC_O_SP; pl.o_sdx = OGB_SPEC_O_SDX;
    pl.o_cache = OGB_SPEC_O_CACHE; pl.o_end = OGB_SPEC_O_END; pl.o_gpert = OGB_SPEC_O_GPERT;
#endif
#endif
#ifdef OGB_SPEC_ZMODE                 // NVRTC build: the zero-stream variant is a compile-time choice too, so the
    zero_mode = OGB_SPEC_ZMODE;       // code of the variants not taken (writer warps, pairs, .cs) is not even compiled
#endif
    const int tid = threadIdx.x, nthr = blockDim.x;
    // %laneid through a volatile asm: the compiler keeps it in a register instead of re-reading
    // SR_TID.X (an S2R costs ~20 cycles) at every use inside the column loop
    int lane;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    const int warp = tid >> 5, nwarps = nthr >> 5;
    OgbWork W;
    W.sbase = smem + pl.o_sbase; W.sc = smem + pl.o_sc; W.scbase = smem + pl.o_scbase;
    W.coef = smem + pl.o_coef; W.prefix = smem + pl.o_prefix; W.pert = smem + pl.o_pert;
    W.pdx = smem + pl.o_pdx; W.px1 = smem + pl.o_px1; W.scpert = smem + pl.o_scpert;
    W.pdlt = smem + pl.o_pdlt; W.pcol = reinterpret_cast<OgbCol*>(smem + pl.o_pcol);
    W.cf = smem + pl.o_cf; W.rterm = smem + pl.o_rterm; W.costp = smem + pl.o_costp;
    W.prdx = smem + pl.o_prdx; W.gpert = smem + pl.o_gpert;
    W.G = pl.G;
    OgbSlot* slots = reinterpret_cast<OgbSlot*>(smem + pl.o_slot);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + pl.o_end);     // [2]
    volatile long* s_next = reinterpret_cast<volatile long*>(smem + pl.o_end + 2);   // next work item [2] (by item parity)

    // ---- once per CTA: problem descriptors and tapes into shared memory
    {
        double* cur = smem + pl.o_cache;
        P.sec = cache_copy(cur, P.sec, (size_t)P.nsec, tid, nthr);
        P.outs = cache_copy(cur, P.outs, (size_t)nouts, tid, nthr);
        P.knots = cache_copy(cur, P.knots, (size_t)P.nknot, tid, nthr);
        P.code = cache_copy(cur, P.code, (size_t)ncode, tid, nthr);
        P.consts = cache_copy(cur, P.consts, (size_t)nconsts, tid, nthr);
    }
    if (tid == 0) { mbar_init(mbar, 1); mbar_init(mbar + 1, 1); }
    __syncthreads();
    const bool fast = NR > 0 && !force_generic;
    for (int s = 0; s < P.nsec; ++s) {
        const OgbSec& S = ogb_sec(P, s);
        for (int t = tid; t < S.nouts; t += nthr) {
            const ogb_out o = P.outs[S.out_off + t];
            OgbSlot si = {0, 0, 0, 0};
            if (t < S.ns) si = OgbSlot{S.rdef + t * S.N, 0, S.N, 1};
            else if (o.kind == OGB_OUT_EQ_POINT || o.kind == OGB_OUT_INEQ_POINT)
                si = OgbSlot{o.row + S.g0 - o.glo, max(0, o.glo - S.g0), min(S.N, o.ghi - S.g0), 0};
            slots[S.out_off + t] = si;
        }
    }
    __syncthreads();
    // Programmatic dependent launch (ogb_eval_fd launches this kernel with programmatic stream serialization right
    // after K1, which signals at its start): everything above -- the problem's own read-only tables -- may run
    // while K1 is still computing D.X; p, D.X, c and J are touched only after K1 has completed and flushed.
    // Without such a launch this is a no-op.
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const int n = P.n, M = P.M, ndx = P.ndx;
    // Writer warps (OGB_OPT_ZERO_MODE bits 2-3; dense FD output only): the last one or two warps of the CTA do
    // nothing but stream the zeros of the CTA's NEXT work item -- its whole region, linearly, 512 bytes per
    // instruction -- while the other (compute) warps run the tapes, assemble c and scatter the non-zeros of the
    // current item, so the SM's store path stays busy through the latency-bound phases.  One full-CTA barrier
    // per item hands over (the next item's index one way, "its zeros are written" the other way); every other
    // barrier of the item loop involves the compute warps only.
    const int nwr = (PACKED_OUT == 0 && with_fd == 1 && (zero_mode & 12) && (nthr >> 5) >= 4) ? ((zero_mode & 8) ? 2 : 1) : 0;
    const int cwarps = (nthr >> 5) - nwr, cnthr = cwarps * 32;        // compute warps / threads
    // work item -> (instance, first column, columns): the instances below `head` in `nchunk` items of pl.group
    // columns, the trailing ones (claimed last) in `tchunk` finer items of pl.tgroup columns (tail refinement)
    const int nchunk = with_fd ? pl.split : 1, tchunk = with_fd ? pl.tsplit : 1;
    const long nheadb = pl.head < B ? pl.head : B;
    const long nhead_items = nheadb * nchunk;
    const long nitems = nhead_items + ((long)B - nheadb) * tchunk;
    auto item_instance = [&](long item) -> long {
        return item < nhead_items ? item / nchunk : nheadb + (item - nhead_items) / tchunk;
    };
    auto item_columns = [&](long item, long b, int& jlo, int& ncols) -> int {     // returns the chunk index
        if (!with_fd) { jlo = 0; ncols = 0; return 0; }
        if (item < nhead_items) {
            const int ch = (int)(item - b * nchunk);
            jlo = ch * pl.group; ncols = min(pl.group, P.n - jlo);
            return ch;
        }
        const int ch = (int)((item - nhead_items) - (b - nheadb) * tchunk);
        jlo = ch * pl.tgroup; ncols = min(pl.tgroup, P.n - jlo);
        return ch;
    };
    const bool fused_dx = DX == nullptr;            // D.X by in-kernel DMMA instead of K1's scratch
    const size_t in_stride = pl.o_sdx - pl.o_sp + ((size_t)(ndx + 2 + 1) & ~(size_t)1);   // doubles per input stage

    // TMA bulk loads of p[b] and D.X[b] into input stage `st` (16-byte aligned body; an odd
    // leading / trailing double is fetched with a plain load by another warp)
    auto stage_inputs = [&](long item, int st) {
        const long b = item_instance(item);
        const double* gp = p + b * n;
        const double* gdx = fused_dx ? nullptr : DX + b * ndx;
        const int hp = (int)((reinterpret_cast<uintptr_t>(gp) >> 3) & 1);
        const int hd = (int)((reinterpret_cast<uintptr_t>(gdx) >> 3) & 1);
        const int bp = (n - hp) & ~1, bd = fused_dx ? 0 : (ndx - hd) & ~1;
        double* sp = smem + pl.o_sp + st * in_stride + hp;       // &sp[hp] is 16-byte aligned
        double* sdx = smem + pl.o_sdx + st * in_stride + hd;
        if (tid == 0) {
            mbar_expect_tx(mbar + st, (uint32_t)(bp + bd) * 8u);
            if (bp) bulk_g2s(sp + hp, gp + hp, (uint32_t)bp * 8u, mbar + st);
            if (bd) bulk_g2s(sdx + hd, gdx + hd, (uint32_t)bd * 8u, mbar + st);
        }
        if (tid == (cnthr > 32 ? 32 : 0)) {      // (a one-warp CTA: thread 0 does both)
            if (hp) sp[0] = gp[0];
            for (int e = hp + bp; e < n; ++e) sp[e] = gp[e];
            if (!fused_dx) {
                if (hd) sdx[0] = gdx[0];
                for (int e = hd + bd; e < ndx; ++e) sdx[e] = gdx[e];
            }
        }
    };

    const int meq = P.meq;

    // Work items are claimed dynamically (one atomic ticket per item; the first item of a CTA is
    // its index) so all CTAs finish within one item of each other.  The ticket for item i+1 is
    // requested at the top of item i and read after the tape phase (its latency is off the critical
    // path); the inputs of item i+1 are then prefetched while item i assembles and streams out.
    auto csync = [&]() {
        if (nwr) asm volatile("bar.sync 1, %0;" ::"r"(cnthr) : "memory");
        else __syncthreads();
    };
    // zeros of one work item's region J[b, jlo .. jlo + ncols, :] by `nw` warps (this thread: lane of warp `w`)
    auto zero_item = [&](long item, int w, int nw) {
        const long zb = item_instance(item);
        int zlo, zn;
        item_columns(item, zb, zlo, zn);
        double* reg = J + (size_t)zb * n * (size_t)M + (size_t)zlo * (size_t)M;
        const size_t len = (size_t)zn * (size_t)M;
        const size_t hj = (reinterpret_cast<uintptr_t>(reg) >> 3) & 1;
        const size_t nv = (len - hj) >> 1;                       // 16-byte pieces
        double2* v = reinterpret_cast<double2*>(reg + hj);
        const double2 z2 = make_double2(0.0, 0.0);
        const size_t step = (size_t)nw * 32;
        size_t i = (size_t)w * 32 + lane;
        for (; i + 3 * step < nv; i += 4 * step) { v[i] = z2; v[i + step] = z2; v[i + 2 * step] = z2; v[i + 3 * step] = z2; }
        for (; i < nv; i += step) v[i] = z2;
        if (w == 0 && lane == 0 && hj) reg[0] = 0.0;
        if (w == 0 && lane == 1 && ((len - hj) & 1)) reg[len - 1] = 0.0;
    };
    if (nwr && (long)blockIdx.x < nitems) zero_item(blockIdx.x, warp, cwarps + nwr);   // the first item: all warps
    if (nwr && warp >= cwarps) {
        // ---- writer warps: zero the next item while the compute warps work on the current one
        unsigned wit = 0;
        for (long item = blockIdx.x; item < nitems; ++wit) {
            __syncthreads();                                     // hand-over: the compute warps published the next item
            const long nx = s_next[wit & 1u];
            if (nx < nitems) zero_item(nx, warp - cwarps, nwr);
            item = nx;
        }
        return;
    }
    if ((long)blockIdx.x < nitems) stage_inputs(blockIdx.x, 0);
    csync();                 // the odd head / tail doubles of the first item are in place
    unsigned it = 0;
    for (long item = blockIdx.x; item < nitems; ++it) {
        const long b = item_instance(item);
        int jlo, ncols;
        const int ch = item_columns(item, b, jlo, ncols);
        const int st = (int)(it & 1u);
        long claimed = 0;
        if (tid == 0)
            claimed = ticket ? (long)gridDim.x + (long)(atomicAdd(ticket, 1ULL) - ticket_base) : item + (long)gridDim.x;

        // ---- phase 1: take this item's inputs (their TMA loads were issued one item ago, so the
        //      copy engine served them long before they are needed)
        {
            const double* gp = p + b * n;
            const double* gdx = fused_dx ? nullptr : DX + b * ndx;
            W.sp = smem + pl.o_sp + st * in_stride + (int)((reinterpret_cast<uintptr_t>(gp) >> 3) & 1);
            W.sdx = smem + pl.o_sdx + st * in_stride + (int)((reinterpret_cast<uintptr_t>(gdx) >> 3) & 1);
        }
        mbar_wait(mbar + st, (it >> 1) & 1u);      // TMA bytes are visible to every thread that waited
        if (with_fd) {               // _check_clip_x (scipy/optimize/_slsqp_py.py:355)
            for (int j = tid; j < n; j += cnthr) {
                const double x = W.sp[j], lo = lb[j], hi = ub[j];
                W.sp[j] = x < lo ? lo : (x > hi ? hi : x);
            }
        }
        csync();

        // ---- phase 2a: D.X of this instance on the FP64 tensor cores (K1 fused in): per phase
        //      OUT[a, i] = sum_l X[a, l] * D[i, l] as m8n8k4 DMMAs, A = the (clipped,
        //      non-dimensional) states in shared memory, B = rows of D through the read-only
        //      path; one (phase, 8 states, 8 nodes) tile per warp at a time; same k order as K1,
        //      so the result is bit-identical to ogb_dx_gemm.
        if (fused_dx) {
            const int qr = lane >> 2, qc = lane & 3;
            int unit0 = 0;
            for (int s = 0; s < P.nsec; ++s) {
                const OgbSec& S = ogb_sec(P, s);
                const int N = S.N, Kp = (N + 3) & ~3;
                const int mt = (S.ns + 7) >> 3, nt = (N + 7) >> 3;
                const double* __restrict__ Dm = P.D + S.doff;
                for (int u = warp - unit0 % cwarps; u < mt * nt; u += cwarps) {
                    if (u < 0) continue;
                    const int a = (u / nt) * 8 + qr, i0 = (u % nt) * 8;
                    const bool av_ok = a < S.ns;
                    const double un = av_ok ? P.ustate[S.us_off + a] : 1.0;
                    const double* xs = W.sp + S.off + a * N;
                    const double* brow = Dm + (i0 + qr) * N;
                    const bool b_ok = i0 + qr < N;
                    double acc0 = 0.0, acc1 = 0.0;
                    for (int l0 = 0; l0 < Kp; l0 += 4) {
                        const int l = l0 + qc;
                        const double av = (av_ok && l < N) ? ogb_nd(xs[l], un) : 0.0;
                        const double bv = (b_ok && l < N) ? __ldg(brow + l) : 0.0;
                        dmma_8x8x4(acc0, acc1, av, bv);
                    }
                    const int ar = (u / nt) * 8 + qr, i = i0 + 2 * qc;
                    if (ar < S.ns) {
                        if (i < N) W.sdx[S.dxoff + ar * N + i] = acc0;
                        if (i + 1 < N) W.sdx[S.dxoff + ar * N + i + 1] = acc1;
                    }
                }
                unit0 += mt * nt;
            }
        }

        // ---- phase 2: tapes -- base nodes, scalar program, one job per Jacobian column
        if (PACKED_OUT == 2) {
            for (int q = tid; q < ogb_njobs(P, ncols); q += cnthr) ogb_job_exact(P, W, q, jlo, ncols);
        } else if (with_fd != 5)        // (probe 5: no tapes, no assembly -- the zero stream alone)
        for (int q = tid; q < ogb_njobs(P, ncols); q += cnthr) ogb_job(P, W, q, jlo, ncols, lb, ub, abs_step);
        if (tid == 0) s_next[it & 1u] = claimed;
        __syncthreads();             // (with writer warps: the one barrier of the item they take part in)
        const long next_item = s_next[it & 1u];
        if (next_item < nitems) stage_inputs(next_item, st ^ 1);      // prefetch one item ahead

        // ---- phase 3: c at the base point and the perturbed cost of every column.  Without a
        //      running cost neither depends on the other, so one barrier covers both.
        if (with_fd != 5) ogb_assemble_base(P, W, tid, cnthr);
        if (!P.has_running && with_fd != 5) {
            if (tid == 0) ogb_assemble_cost(P, W);
            for (int cl = tid; cl < ncols; cl += cnthr) {
                if (PACKED_OUT == 2) ogb_cost_column_exact(P, W, cl); else ogb_cost_column(P, W, cl);
            }
        }
        csync();
        if (P.has_running) {
            if (tid == 0) ogb_assemble_cost(P, W);
            csync();
            for (int cl = tid; cl < ncols; cl += cnthr) {
                if (PACKED_OUT == 2) ogb_cost_column_exact(P, W, cl); else ogb_cost_column(P, W, cl);
            }
            csync();
        }
        if (ch == 0)
            for (int r = tid; r < M; r += cnthr) c[b * M + r] = W.sc[r];

        // ---- phase 4: Jacobian columns, one warp per column (columns warp, warp + cwarps, ...),
        //      no block barrier and no staging: the warp streams the column's zeros to HBM with
        //      16-byte stores, then (ordered by __syncwarp) overwrites the few rows that can be
        //      non-zero.  The overwrites hit sectors still resident in L2, so DRAM sees each
        //      sector once.
        //      The loop is written for a short dependent chain (the kernel runs at ~0.45 IPC, bound
        //      by instruction latency, not by HBM): every operand of the column is loaded first
        //      (shared memory through constant offsets, D^T through the read-only path), the zero
        //      stream is issued while those loads fly, and the column pointer is carried from
        //      iteration to iteration instead of being rebuilt from (instance, column) per store.
        // zero_mode (experiments, OGB_OPT_ZERO_MODE): bit 0 = streaming (st.global.cs) zero stores (measured slower);
        // bits 2-3 = dedicated writer warps (see below)
        const bool zero_cs = (zero_mode & 1) != 0;
        auto columns = [&](auto packed_tag) {
            constexpr bool PACKED = decltype(packed_tag)::value;
            constexpr int NRA = NR > 0 ? NR : 1;
            // dense: the item's first column J[b, jlo, :]; packed: the instance's values [nnz]
            double* const gbase = PACKED ? J + (size_t)b * (size_t)P.nnz
                                         : J + (size_t)b * n * (size_t)M + (size_t)jlo * (size_t)M;
            // Work units of the column phase.  Normally one column each.  With an odd number of rows M a column
            // starts 8 bytes off the 16-byte grid every other time, and its zero stream needs an 8-byte store at
            // each end (partial sectors, shared with the neighbour column: the fill microbenchmarks lose 13 % to
            // that).  Pair mode (dense output, M odd, zero_mode bit 4): a warp takes two ADJACENT columns that start
            // on the 16-byte grid together and zeroes them as ONE aligned span of 2 M doubles, then writes the
            // non-zeros of both; only a leading / trailing single column of the item keeps an 8-byte end.
            const bool pairs = !PACKED && (M & 1) && (zero_mode & 16) && with_fd == 1 && !nwr;
            const int lead = pairs ? (int)((reinterpret_cast<uintptr_t>(gbase) >> 3) & 1) : 0;   // item starts off-grid
            const int npairs = pairs ? (ncols - lead) / 2 : 0;
            const int nunits = pairs ? lead + npairs + ((ncols - lead) & 1) : ncols;
            int cur_key = -1, slot_sec = -1;
            double r_sdx[NRA], r_cf[NRA], r_sc[NRA];
            OgbSlot si = {0, 0, 0, 0};               // where this lane's output slot lands (per phase)
            const double* const s_pdx = smem + pl.o_pdx;
            const double* const s_prdx = smem + pl.o_prdx;
            const double* const s_pdlt = smem + pl.o_pdlt;
            const double* const s_pert = smem + pl.o_pert;
            const double* const s_sc = smem + pl.o_sc;
            const double* const s_cf = smem + pl.o_cf;
            const double* const s_coef = smem + pl.o_coef;
            const int4* const s_pcol = reinterpret_cast<const int4*>(smem + pl.o_pcol);
            const double* const s_sdx = W.sdx;       // (the input stage alternates between items)
            for (int u = warp; u < nunits; u += cwarps) {
              // the unit's first column, how many columns it has, and the length of the zero span it starts
              int c0 = u, ucols = 1;
              if (pairs) {
                  if (u < lead) c0 = 0;
                  else if (u < lead + npairs) { c0 = lead + 2 * (u - lead); ucols = 2; }
                  else c0 = ncols - 1;
              }
              for (int t = 0; t < ucols; ++t) {
                const int cc = c0 + t;
                const int zlen = t == 0 ? ucols * M : 0;             // doubles to zero from this column's start
                double* gdst = PACKED ? gbase : gbase + (size_t)cc * (size_t)M;
                const int* pm = PACKED ? P.pmap + (size_t)(jlo + cc) * (size_t)M : nullptr;   // row -> packed entry
                asm volatile("" : "+l"(gdst));       // keep the column pointer in registers ...
                __builtin_assume(__isGlobal(gdst));  // ... and its stores in the global space (STG, not ST)
                // (A) operands
                const int4 cdv = s_pcol[cc];
                const OgbCol cd = {cdv.x, cdv.y, cdv.z, cdv.w};
                const double dx = s_pdx[cc], rdx = s_prdx[cc];
                const bool fcol = fast && cd.sec >= 0 && !(P.any_global && P.gcol_of[jlo + cc] >= 0);
                double dlt = 0.0, pv = 0.0, dkk = 0.0, coef = 0.0;
                double dtv[NRA];
#pragma unroll
                for (int r = 0; r < NRA; ++r) dtv[r] = 0.0;
                int a = -1;
                int pos_blk = 0, pos_slot = -1;      // packed: entry of the state block's first row / of this lane's slot row
                if (fcol) {
                    const OgbSec& S = ogb_sec(P, cd.sec);
                    dlt = s_pdlt[cc];
                    coef = s_coef[3 * cd.sec];
                    if (cd.blk < S.ns) {
                        a = cd.blk;
                        const double* __restrict__ Dt = P.Dt + S.doff + cd.k * S.N;
#pragma unroll
                        for (int r = 0; r < NRA; ++r) {
                            const int i = lane + 32 * r;
                            if (i < S.N) dtv[r] = __ldg(Dt + i);
                        }
                        dkk = __ldg(Dt + cd.k);
                        // the N rows of state block a are consecutive entries of the column (row k is the
                        // node row of slot a, the others are the D-block)
                        if (PACKED) pos_blk = __ldg(pm + S.rdef + a * S.N);
                    }
                    if (lane < S.nouts) pv = s_pert[lane * W.G + cc];
                }
                // (B) zeros: 16-byte aligned body, an odd first / last double on its own
                //     (with_fd == 2: structure probe -- only the overwrites below land, on a
                //      sentinel-filled J; see ogb_jac_pattern.  3 / 4 / 5: timing probes)
                if (!PACKED && with_fd != 2 && with_fd != 4 && !nwr && zlen) {
                    const unsigned hj = (unsigned)((reinterpret_cast<uintptr_t>(gdst) >> 3) & 1);
                    const unsigned nbytes = ((unsigned)(zlen - hj) & ~1u) * 8u;
                    char* g = reinterpret_cast<char*>(gdst + hj) + lane * 16;
                    const double2 z2 = make_double2(0.0, 0.0);
                    unsigned left = nbytes;                      // bytes not yet covered by the warp
                    if (zero_cs) {
                        for (; left >= 512u; left -= 512u, g += 512) __stcs(reinterpret_cast<double2*>(g), z2);
                    }
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
                    if (lane == 0 && hj) gdst[0] = 0.0;
                    if (lane == 1 && ((zlen - hj) & 1)) gdst[zlen - 1] = 0.0;
                }
                __syncwarp();
                if (!PACKED && with_fd >= 3) continue;   // timing probes: zero stream only / no column output
                // (C) the rows that can be non-zero
                const int j = jlo + cc;
                if (fcol) {
                    const OgbSec& S = ogb_sec(P, cd.sec);
                    const int N = S.N, k = cd.k;
                    if (cd.sec != slot_sec) {                    // new phase: this lane's slot record
                        slot_sec = cd.sec;
                        si = lane < S.nouts ? slots[S.out_off + lane] : OgbSlot{0, 0, 0, 0};
                    }
                    const bool slot_hit = k >= si.klo && k < si.khi;
                    if (PACKED && slot_hit) pos_slot = __ldg(pm + si.rbase + k);
                    if (a >= 0) {
                        const int key = cd.sec * 1024 + a;
                        if (key != cur_key) {                    // new state block: reload row constants
                            cur_key = key;
#pragma unroll
                            for (int r = 0; r < NRA; ++r) {
                                const int i = lane + 32 * r;
                                if (i < N) {
                                    r_sdx[r] = s_sdx[S.dxoff + a * N + i];
                                    r_cf[r] = s_cf[S.dxoff + a * N + i];
                                    r_sc[r] = s_sc[S.rdef + a * N + i];
                                }
                            }
                        }
                        double* crow = PACKED ? gdst + (pos_blk + lane) : gdst + (S.rdef + a * N + lane);
#pragma unroll
                        for (int r = 0; r < NRA; ++r) {
                            const int i = lane + 32 * r;
                            if (i < N && i != k) {
                                const double cp = (r_sdx[r] + dtv[r] * dlt) - r_cf[r];
                                crow[32 * r] = ogb_fd_div(cp - r_sc[r], dx, rdx);
                            }
                        }
                    }
                    if (slot_hit) {                              // slot t = lane (slots 0..31)
                        double cp = pv;
                        const int r = si.rbase + k;
                        if (si.isdyn) {
                            double dxp = s_sdx[S.dxoff + lane * N + k];
                            if (lane == a) dxp = dxp + dkk * dlt;
                            cp = dxp - coef * cp;
                        }
                        gdst[PACKED ? pos_slot : r] = ogb_fd_div(cp - s_sc[r], dx, rdx);
                    }
                    for (int t = lane + 32; t < S.nouts; t += 32) {   // (rare) more than 32 output slots
                        const OgbSlot s2 = slots[S.out_off + t];
                        if (k >= s2.klo && k < s2.khi) {
                            double cp = s_pert[t * W.G + cc];
                            const int r = s2.rbase + k;
                            if (s2.isdyn) {
                                double dxp = s_sdx[S.dxoff + t * N + k];
                                if (t == a) dxp = dxp + dkk * dlt;
                                cp = dxp - coef * cp;
                            }
                            gdst[PACKED ? __ldg(pm + r) : r] = ogb_fd_div(cp - s_sc[r], dx, rdx);
                        }
                    }
                    if (PACKED) {
                        const OgbColPacked col{gdst, pm};
                        if (P.nknot && (k == 0 || k == N - 1)) ogb_scatter_knots(P, W, j, W.px1[cc], dx, rdx, col, lane, 32);
                        if (cd.pick >= 0 || P.has_running) ogb_scatter_scalar_cost(P, W, cd, cc, dx, rdx, col, lane, 32);
                    } else {
                        const OgbColOut col{gdst, gdst + meq, meq};
                        if (P.nknot && (k == 0 || k == N - 1)) ogb_scatter_knots(P, W, j, W.px1[cc], dx, rdx, col, lane, 32);
                        if (cd.pick >= 0 || P.has_running) ogb_scatter_scalar_cost(P, W, cd, cc, dx, rdx, col, lane, 32);
                    }
                } else if (PACKED) {
                    ogb_scatter_column(P, W, j, cc, OgbColPacked{gdst, pm}, lane, 32);
                } else {
                    ogb_scatter_column(P, W, j, cc, OgbColOut{gdst, gdst + meq, meq}, lane, 32);
                }
              }
            }
        };
        if (PACKED_OUT == 2) {
            // exact mode: one warp per column, every structural non-zero straight into the packed values.
            // Fast path (like the FD one): the D-block of a state column is one contiguous run of packed entries
            // holding a column of D, lane t owns output slot t of the node program.
            double* vb = J + (size_t)b * (size_t)P.nnz;
            const double* const s_pert = smem + pl.o_pert;
            const double* const s_coef = smem + pl.o_coef;
            const int4* const s_pcol = reinterpret_cast<const int4*>(smem + pl.o_pcol);
            int slot_sec = -1;
            OgbSlot si = {0, 0, 0, 0};
            for (int cc = warp; cc < ncols; cc += cwarps) {
                const int j = jlo + cc;
                const int* pm = P.pmap + (size_t)j * (size_t)M;
                const int4 cdv = s_pcol[cc];
                const OgbCol cd = {cdv.x, cdv.y, cdv.z, cdv.w};
                const OgbColPacked col{vb, pm};
                if (!(fast && cd.sec >= 0) || (P.any_global && P.gcol_of[j] >= 0)) {
                    ogb_scatter_column_exact(P, W, j, cc, col, lane, 32);
                    continue;
                }
                const OgbSec& S = ogb_sec(P, cd.sec);
                const int N = S.N, k = cd.k, a = cd.blk < S.ns ? cd.blk : -1;
                if (cd.sec != slot_sec) {
                    slot_sec = cd.sec;
                    si = lane < S.nouts ? slots[S.out_off + lane] : OgbSlot{0, 0, 0, 0};
                }
                const double* __restrict__ Dt = P.Dt + S.doff + k * N;
                double dkk = 0.0;
                if (a >= 0) {
                    const int pos_blk = __ldg(pm + S.rdef + a * N);
                    dkk = __ldg(Dt + k);
                    for (int i = lane; i < N; i += 32)
                        if (i != k) vb[pos_blk + i] = __ldg(Dt + i);
                }
                const double coef = s_coef[3 * cd.sec];
                if (k >= si.klo && k < si.khi) {
                    const double t = s_pert[lane * W.G + cc];
                    vb[__ldg(pm + si.rbase + k)] = si.isdyn ? ((lane == a ? dkk : 0.0) - coef * t) : t;
                }
                for (int t2 = lane + 32; t2 < S.nouts; t2 += 32) {
                    const OgbSlot s2 = slots[S.out_off + t2];
                    if (k >= s2.klo && k < s2.khi) {
                        const double t = s_pert[t2 * W.G + cc];
                        vb[__ldg(pm + s2.rbase + k)] = s2.isdyn ? ((t2 == a ? dkk : 0.0) - coef * t) : t;
                    }
                }
                if (P.nknot && (k == 0 || k == N - 1)) ogb_scatter_knots_exact(P, j, col, lane, 32);
                if (cd.pick >= 0 || P.has_running) ogb_scatter_scalar_cost_exact(P, W, cd, cc, col, lane, 32);
            }
        } else {
            columns(OgbBool<PACKED_OUT != 0>{});
        }
        csync();             // all warps are done reading this item's staging
        item = next_item;
    }
}

