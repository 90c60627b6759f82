/* ogb200.h -- C ABI of libogb200.so, the B200 (sm_100a) engine behind the
 * `OpenGoddard.optimize` facade.
 *
 * The reference (istellartech/OpenGoddard) is pure Python and has no FFI; its
 * boundary for the hot path is the set of Python closures `Problem.solve` hands to
 * SciPy (reference OpenGoddard/optimize.py:670-749).  Each entry point below cites
 * the reference code it replaces.  All `double*` arguments are CUDA DEVICE pointers
 * (e.g. torch.Tensor.data_ptr()) unless the name ends in `_h`; the caller owns every
 * buffer; every launch is enqueued on the given stream and does not synchronise.
 * Return value: 0 = ok, negative = error (text via ogb_last_error()).
 *
 * One *eval* (the benchmark unit) = the stacked vector c = [c_eq ; c_ineq ; cost]
 * at one decision vector p plus its dense forward-difference Jacobian with respect
 * to all nvars variables -- what SLSQP asks for in mode -1
 * (scipy/optimize/_slsqp_py.py:532-534).
 */
#ifndef OGB200_H
#define OGB200_H

#ifndef __CUDACC_RTC__
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define OGB_VERSION 200

/* ---- expression tapes -----------------------------------------------------
 * User callbacks (dynamics / equality / inequality / cost / running_cost; reference
 * optimize.py:674,685,703,706,727) are traced once on the host into straight-line
 * register programs.  One instruction = one 64-bit word:
 *     op[63:56]  dst[55:42]  a[41:28]  b[27:14]  c[13:0]
 * A *node program* runs once per LGL node: OGB_LDP reads block `a` (state or control
 * number within the phase) of the decision vector at that node; operands beyond the phase's
 * blocks address first the program's per-node constant vectors (nodec_h: e.g. prob.time[s],
 * reference optimize.py:786-791), then its global variables (globals_h: single decision
 * variables read at every node -- final times, optimize.py:349-360, non-autonomous dynamics; picked state /
 * control elements such as x[0] or the operands of sum(x) inside `dynamics`).  The *scalar
 * program* runs once per decision vector: OGB_LDP reads variable `a` of p.          */
enum ogb_opcode {
    OGB_NOP = 0,
    OGB_LDP = 1,   /* r[dst] = input(a)                                   */
    OGB_LDC = 2,   /* r[dst] = consts[a]                                  */
    OGB_OUT = 3,   /* output slot `dst` = r[a]                            */
    OGB_ADD = 4, OGB_SUB = 5, OGB_MUL = 6, OGB_DIV = 7,
    OGB_POW = 8, OGB_MIN = 9, OGB_MAX = 10, OGB_ATAN2 = 11,
    OGB_LT = 12, OGB_LE = 13, OGB_GT = 14, OGB_GE = 15, OGB_EQ = 16, OGB_NE = 17,
    OGB_SEL = 18,  /* r[dst] = r[a] != 0 ? r[b] : r[c]                    */
    OGB_NEG = 19, OGB_SQRT = 20, OGB_EXP = 21, OGB_LOG = 22, OGB_SIN = 23,
    OGB_COS = 24, OGB_TAN = 25, OGB_ABS = 26, OGB_SQUARE = 27, OGB_RECIP = 28,
    OGB_ASIN = 29, OGB_ACOS = 30, OGB_ATAN = 31, OGB_SINH = 32, OGB_COSH = 33,
    OGB_TANH = 34, OGB_LOG10 = 35, OGB_SIGN = 36, OGB_FLOOR = 37, OGB_CEIL = 38,
    OGB_AND = 39, OGB_OR = 40, OGB_NOT = 41,
    OGB_INTERP = 42, /* r[dst] = table b evaluated at r[a] (scipy.interpolate.interp1d, linear)  */
    OGB_OP_COUNT = 43
};
#define OGB_MAX_REG 96      /* registers per tape (host compiler enforces)   */
#define OGB_MAX_FIELD 16383 /* 14-bit operand fields                         */

/* what an output slot of a tape feeds */
enum ogb_out_kind {
    OGB_OUT_DYN = 0,        /* node prog: f_a at the node, already scaled by unit_time/unit_state
                               (reference Dynamics.__call__, optimize.py:1122-1127); row = a   */
    OGB_OUT_EQ_POINT = 1,   /* node prog: one user equality row per node; row = first row of the
                               block (relative to the user-equality rows), nodes [glo, ghi)     */
    OGB_OUT_INEQ_POINT = 2, /* same for user inequality rows                                    */
    OGB_OUT_RUNNING = 3,    /* node prog: running-cost integrand at the node (optimize.py:706)  */
    OGB_OUT_EQ_SCALAR = 4,  /* scalar prog: user equality row `row`                             */
    OGB_OUT_INEQ_SCALAR = 5,/* scalar prog: user inequality row `row`                           */
    OGB_OUT_COST = 6        /* scalar prog: non-integrated cost (optimize.py:703)               */
};

typedef struct ogb_out {
    int32_t kind;
    int32_t row;
    int32_t glo, ghi;       /* global node range (over concatenated phases) for *_POINT kinds  */
} ogb_out;

/* A 1-D linear lookup table traced from a `scipy.interpolate.interp1d(x, y)` object called
 * inside a callback (reference examples/11_Polar_TSTO_Taiki.py:21-27,94-98).  variant 0 = the
 * np.interp formula SciPy uses for float tables that do not extrapolate
 * (slope*(x - x_lo) + y_lo), variant 1 = SciPy's two-term formula
 * ((x-x_lo)/(x_hi-x_lo)*y_hi + (x_hi-x)/(x_hi-x_lo)*y_lo, used with fill_value="extrapolate").
 * Without extrapolation x below / above the table yields fill_below / fill_above (NaN where the
 * reference would raise because bounds_error is set).                                        */
typedef struct ogb_table {
    int32_t off, len;        /* slice of table_x_h / table_y_h (x ascending)                    */
    int32_t variant, extrapolate;
    double fill_below, fill_above;
} ogb_table;

typedef struct ogb_program {
    const uint64_t* code_h;   int32_t ncode;
    const double*   consts_h; int32_t nconsts;
    const ogb_out*  outs_h;   int32_t nouts;   /* slot i of OGB_OUT <-> outs_h[i] */
    int32_t nreg;
    const double*   nodec_h;  int32_t n_nodec; /* node programs: n_nodec vectors of `nodes` doubles (operand nb + i) */
    const int32_t*  globals_h; int32_t nglobals; /* node programs: decision-variable indices (operand nb + n_nodec + i);      
                                                  the Jacobian columns of these variables are dense in the phase  */
} ogb_program;

/* Everything `Problem.__init__` + the unit setters + the traced callbacks define
 * (reference optimize.py:759-823, :579-639).  All pointers are HOST pointers.      */
typedef struct ogb_problem_desc {
    int32_t nsec;                 /* number_of_section                                */
    const int32_t* nodes_h;       /* [nsec]                                           */
    const int32_t* nstates_h;     /* [nsec]                                           */
    const int32_t* ncontrols_h;   /* [nsec]                                           */
    const double*  unit_states_h; /* unit_states, phases concatenated [sum nstates]   */
    double unit_time;
    double t0;                    /* prob.t0 as `time_start(0)` returns it (:343-344) */
    const uint8_t* knot_smooth_h; /* [nsec-1] knot_states_smooth                      */
    int32_t meq_user;             /* rows returned by the user equality               */
    int32_t mineq_user;           /* rows returned by the user inequality             */
    int32_t has_running_cost;
    const ogb_program* node_prog_h;   /* [nsec] one node program per phase            */
    const ogb_program* scalar_prog_h; /* exactly one (may have ncode == 0 outs == 0)  */
    int32_t ntables;                  /* lookup tables referenced by OGB_INTERP       */
    const ogb_table* tables_h;        /* [ntables]                                    */
    const double* table_x_h;          /* concatenated abscissae                       */
    const double* table_y_h;          /* concatenated ordinates                       */
    const double* unit_controls_h;    /* unit_controls, phases concatenated [sum ncontrols]; NULL = all 1.
                                         Only the guess / trajectory kernels use it (inside the traced
                                         callbacks the units are already part of the tapes)            */
} ogb_problem_desc;

typedef struct ogb_problem_info {
    int32_t nvars;      /* number_of_variables (optimize.py:781)                      */
    int32_t meq;        /* user eq rows + collocation defects + knot rows             */
    int32_t mineq;      /* user inequality rows                                       */
    int32_t nrows;      /* meq + mineq + 1 (the last row carries cost / grad cost)    */
    int32_t ndx;        /* doubles of D.X per instance = sum_s nstates_s * nodes_s    */
    int32_t total_nodes;
    int32_t tile_cols;  /* warps per CTA of the sweep kernel (one column pipeline each) */
    int32_t group_cols; /* Jacobian columns per work item                             */
    int32_t smem_bytes; /* dynamic shared memory of the sweep kernel                  */
    int32_t ctas_per_sm;
    int32_t jit;        /* 1: the tapes were compiled into the sweep kernel with NVRTC      */
    int32_t nnz;        /* structural non-zeros of one instance's J, -1 until the pattern is known   */
    int64_t launches;   /* kernels launched through this handle so far (counted where they are launched) */
} ogb_problem_info;

const char* ogb_last_error(void);
int ogb_version(void);

/* LGL nodes, weights, differentiation matrix (row-major N x N).
 * Replaces Problem._nodes_LGL / _weight_LGL / _differentiation_matrix_LGL
 * (reference optimize.py:183-213).  `ogb_lgl_build` writes DEVICE memory from a
 * kernel; `ogb_lgl_build_host` runs the same code on the host for the facade's
 * constructor attributes (prob.tau / w / D / time; optimize.py:786-791).          */
int ogb_lgl_build(int N, double* tau, double* w, double* D, void* stream);
int ogb_lgl_build_host(int N, double* tau_h, double* w_h, double* D_h);

/* Problem handle on the current CUDA device.  Replaces the layout tables
 * (_make_param_division, optimize.py:237-245) and owns D per phase on the device. */
void* ogb_problem_create(const ogb_problem_desc* desc);
void  ogb_problem_destroy(void* prob);
int   ogb_problem_info_get(void* prob, ogb_problem_info* out);

/* Tuning / verification knobs (not needed for normal use). */
enum ogb_option {
    OGB_OPT_GENERIC_COLUMNS = 0, /* 1: produce Jacobian columns with the generic per-row code instead of
                                    the register-cached fast path (results must be bit-identical)     */
    OGB_OPT_THREADS = 1,         /* CTA size of the sweep kernel: a multiple of 32 up to 256 (up to 512 for the
                                    NVRTC-specialised kernel only); DeviceProblem.autotune tries 256 and 384  */
    OGB_OPT_JIT = 2,             /* 1: run the NVRTC-specialised sweep kernel (tapes compiled to device
                                    code), 0: the ahead-of-time kernel with the tape interpreter       */
    OGB_OPT_GRID_CAP = 3,        /* cap on the persistent grid (0 = SM count x resident CTAs)         */
    OGB_OPT_DYNAMIC_ITEMS = 5,   /* 1 (default): persistent CTAs claim work items with an atomic ticket;
                                    0: static round-robin assignment                                    */
    OGB_OPT_GROUP_COLS = 6,      /* cap on Jacobian columns per work item (default 256); an instance is
                                    split into ceil(nvars / cap) items                                 */
    OGB_OPT_AUTO_SPLIT = 7,      /* 1 (default): batches with fewer than ~6 instances per resident CTA are cut
                                    into more, smaller work items (bit-identical results, better balance)  */
    OGB_OPT_PROBE_MODE = 8,      /* timing probes only (J is NOT a Jacobian afterwards): 2 = no zero stream,
                                    3 = zero stream only, 4 = no column output at all; 0 = normal             */
    OGB_OPT_SPLIT = 9,           /* how ogb_eval_fd produces the dense J: 0 = the fused sweep kernel (one launch after
                                    K1), 1 = the split pipeline (K1 -> K2a packed sweep on an internal stream, K2b
                                    densify on the caller's stream, chunk by chunk), -1 (default) = automatic, which
                                    currently means fused: it measured faster at every batch size (profiles/README.md).
                                    Results are bit-identical either way.                                         */
    OGB_OPT_SPLIT_CHUNK = 10,    /* instances per chunk of the split pipeline (0 = auto, ~192 MB of dense J)        */
    OGB_OPT_DENSE_STREAMING = 11,/* 1 (default): K2b writes its zeros with st.global.cs (evict-first)              */
    OGB_OPT_ZERO_MODE = 12,      /* how the fused kernel writes its zeros (results unchanged): 0 = every warp streams the
                                    zeros of its own column, then overwrites the non-zeros; bit 0 = with st.global.cs
                                    (measured slower); 4 / 8 = one / two dedicated writer warps per CTA stream the
                                    zeros of the CTA's next work item while the other warps compute (measured slower);
                                    16 = with an odd number of rows a warp takes two adjacent columns and zeroes them
                                    as one 16-byte-aligned span (no 8-byte stores at the column ends)                 */
    OGB_OPT_GEMM_UNIT = 13,      /* K1 form: 0 (default) = the latency-organised kernel (compile-time strides, cp.async
                                    staging, hoisted unit division; phases of <= 128 nodes); 8 = the round-1 kernel, an
                                    8-row tile computes whole rows of D.X; 2 = its (8-row tile, 16 output nodes) units
                                    (experiment, measured slower).  All three give the same bits                     */
    OGB_OPT_PDL = 14,            /* 1 (default): the sweep kernel is launched behind K1 with programmatic stream
                                    serialization (its launch and per-CTA prologue overlap K1; results unchanged)  */
    OGB_OPT_TAIL_REFINE = 15,    /* tail refinement of the dense FD sweep: the last instances of a large batch (this
                                    many per cent of one wave of work items) are cut into items of >= 64 columns so
                                    the persistent CTAs finish together; 0 = off; -1 (default) = 100 for problems with
                                    light node programs (<= 64 tape operations), off for heavy ones.  Results unchanged */
    OGB_OPT_FUSED_DX = 4         /* 0: K1 ogb_dx_gemm writes the D.X scratch, then the sweep (two launches); 1: the sweep
                                    kernel computes D.X itself with in-kernel DMMAs (one launch; bit-identical);
                                    -1 (default): automatic -- one launch for batches of at most half a wave of CTAs
                                    (23-31 % faster at B <= 128), two launches above (8-27 % faster from B = 512)    */
};
int ogb_problem_set_option(void* prob, int key, int value);

/* Generate and NVRTC-compile the specialised sweep kernel for `desc` without loading it (works
 * without a GPU).  Returns the cubin size (> 0) and copies the generated source into `log`, or a
 * negative value and the compiler log.                                                      */
int ogb_jit_check(const ogb_problem_desc* desc, char* log, int log_cap);
/* ... the same for one of the three kernel variants: 0 = dense J / c only (what ogb_jit_check builds),
 * 1 = packed FD output (ogb_eval_sparse), 2 = exact mode (ogb_eval_exact).  Each variant is compiled
 * when it is first used.                                                                       */
int ogb_jit_check_variant(const ogb_problem_desc* desc, int variant, char* log, int log_cap);

/* Scratch the caller must provide to ogb_eval / ogb_eval_fd for a batch of B.     */
size_t ogb_workspace_bytes(void* prob, int B);

/* K1: D.X for every phase and state of every instance as a batched FP64
 * tensor-core GEMM.  Replaces `D[i].dot(state_temp)` (reference optimize.py:680-682).
 * p [B, nvars] -> DX [B, ndx] (phase-major, then state-major, then node).  With lb / ub
 * (both or neither) p is clipped into the bounds first, as ogb_eval_fd does.          */
int ogb_dx_gemm(void* prob, const double* p, const double* lb, const double* ub, int B, double* DX,
                void* stream);

/* K2 alone (ogb_eval / ogb_eval_fd = ogb_dx_gemm + ogb_sweep on one stream), exposed so the
 * dominant kernel can be timed and profiled by itself: DX must come from ogb_dx_gemm on the
 * same p and bounds, or NULL to let the kernel compute D.X itself (OGB_OPT_FUSED_DX).
 * J == NULL: constraint vector only (no clipping).                                      */
int ogb_sweep(void* prob, const double* p, const double* DX, const double* lb, const double* ub,
              double abs_step, int B, double* c, double* J, void* stream);

/* c = [c_eq ; c_ineq ; cost] at every p[b] (no clipping).  Replaces the
 * `for_solver(equality_add)`, `for_solver(inequality)` and `for_solver(cost_add)`
 * closures (reference optimize.py:670-715).  c [B, nrows].                        */
int ogb_eval(void* prob, const double* p, int B, double* c, void* work, void* stream);

/* One eval per instance: c at x = clip(p[b], lb, ub) and the dense 2-point forward
 * difference Jacobian of c, bound-adjusted steps, dx = (x+h)-x as divisor.
 * Replaces SciPy's cjac -> approx_derivative -> _dense_difference on the reference's
 * closures (scipy/optimize/_slsqp_py.py:353-367, _numdiff.py:14-90,582-600,683-712).
 * J [B, nvars, nrows]: J[b, j, :] is the column of variable j (contiguous), i.e.
 * the Fortran-ordered (nrows, nvars) matrix SLSQP consumes.  lb / ub [nvars] use
 * +-inf for "no bound".                                                           */
int ogb_eval_fd(void* prob, const double* p, const double* lb, const double* ub,
                double abs_step, int B, double* c, double* J, void* work, void* stream);

/* ---- sparse evaluation -------------------------------------------------------------------
 * ogb_eval_sparse = ogb_eval_fd without the dense Jacobian: c [B, nrows] and vals [B, nnz], the
 * structurally non-zero entries of every instance's FD Jacobian in the ogb_jac_pattern layout
 * (K1 + the sweep kernel writing packed output; bit-identical to the same entries of ogb_eval_fd's
 * J).  ogb_densify (kernel K2b) expands packed values into the dense J [B, nvars, nrows], zeros
 * included; ogb_eval_fd's split pipeline is the two chained chunk by chunk.  Replaces the same
 * reference code as ogb_eval_fd (scipy/optimize/_slsqp_py.py:353-367 on optimize.py:670-715).      */
int ogb_eval_sparse(void* prob, const double* p, const double* lb, const double* ub, double abs_step,
                    int B, double* c, double* vals, void* work, void* stream);
int ogb_densify(void* prob, const double* vals, int B, double* J, void* stream);

/* ---- initial guesses and post-processing on the device ------------------
 * Batched counterparts of the reference's one-instance host helpers, so that large multi-start batches
 * are generated, perturbed and unpacked without a host round trip.
 *
 * ogb_guess_fill: Guess.zeros / constant / linear / cubic (reference optimize.py:883-956) evaluated on the
 * problem's LGL time nodes and stored the way set_states / set_controls(_all_section) store them (value /
 * unit, :377-440).  Spec i writes block `blk` (states first, then controls) of phase `sec`, or of every
 * phase over the concatenated time axis when sec = -1 (the *_all_section calls the shipped examples make:
 * examples/04_Goddard_0knot.py:115-140); instance b takes its parameters from params[b, i, 0..3]:
 * constant c | linear (y0, yf) | cubic (y0, y'0, yf, y'f).  time_nodes [total_nodes]: prob.time_all_section.
 * tfinal (may be NULL) [B, nsec]: final times to store (set_time_final, :437-440).  Entries of P no spec
 * covers are left untouched.                                                                          */
enum ogb_guess_kind { OGB_GUESS_ZEROS = 0, OGB_GUESS_CONSTANT = 1, OGB_GUESS_LINEAR = 2, OGB_GUESS_CUBIC = 3 };
typedef struct ogb_guess_spec {
    int32_t sec, blk, kind, pad;
} ogb_guess_spec;
int ogb_guess_fill(void* prob, const ogb_guess_spec* specs_h, int nspec, const double* params /*[B, nspec, 4]*/,
                   const double* time_nodes /*[total_nodes]*/, const double* tfinal /*[B, nsec] or NULL*/, int B,
                   double* P /*[B, nvars]*/, void* stream);

/* ogb_jitter: the multi-start perturbation of a batch of decision vectors, in place:
 *   state / control entries  p <- p * (1 + rel_x * z),  z ~ N(0, 1);   final times  p <- p * (1 + rel_t * u),
 *   u ~ U[-1, 1);   then clipped into [lb, ub] (either may be NULL).
 * Counter-based generator (Philox4x32-10, key = seed, counter = (variable, 0, first_instance + b, 0); Box-Muller
 * on two 53-bit uniforms), so any sub-range of instances is reproducible on any number of GPUs; restated
 * in numpy by the test oracle (og_rng) for the parity test.                                                  */
int ogb_jitter(void* prob, double* P, int B, uint64_t seed, int64_t first_instance, double rel_x, double rel_t,
               const double* lb, const double* ub, void* stream);

/* ogb_trajectories: what Problem.time_update / states_all_section / controls_all_section / to_csv assemble
 * per instance (reference optimize.py:518-531, :286-331, :844-863): out [B, total_nodes, 1 + ns + nc] holds per
 * node the dimensional time ((t_{s+1} - t_s) / 2 * tau + (t_{s+1} + t_s) / 2 with t = [0, final times], as the
 * reference assumes t0 = 0 there), states and controls (p * unit); ns / nc are those of phase 0, as in to_csv. */
int ogb_trajectories(void* prob, const double* P, int B, double* out, void* stream);

/* ---- exact Jacobian mode (opt-in) ------------------------------------
 * The same rows and columns with derivatives instead of difference quotients: the collocation block is
 * analytic (D_p (x) I: d defect(a, i) / d x(a, k) = D[i, k]; reference optimize.py:677-696), everything that
 * goes through a user callback is differentiated in forward mode -- the traced tapes run in dual
 * arithmetic, one tangent per perturbed block -- at x = clip(p, lb, ub).  Output: c [B, nrows] (bit-identical
 * to ogb_eval_fd's) and vals [B, nnz] in the ogb_jac_pattern layout (ogb_densify gives the dense matrix).
 * This is NOT what the reference computes (SciPy's forward differences carry a truncation error of up to
 * ~1e-5 of the row maximum and rounding noise of relative size 1 on small entries); it is what SLSQP would
 * ideally be given.                                                                              */
int ogb_eval_exact(void* prob, const double* p, const double* lb, const double* ub, int B, double* c,
                   double* vals, void* work, void* stream);

/* ---- packed Jacobian transport ---------------------------------------------------------
 * The FD Jacobian is structurally sparse: a perturbed state moves its own defect rows and the
 * rows living at its node (eq J 7 758 / 31 155, ineq J 301 / 60 501
 * non-zeros at Goddard-50).  Which entries of an instance's J [nvars, nrows] can be non-zero is a
 * property of the problem, not of p.  ogb_jac_pattern returns their count and (if lin_h != NULL,
 * cap >= count) their ascending linear indices j * nrows + r; ogb_pack gathers them from a dense
 * device J [B, nvars, nrows] into vals [B, nnz] (kernel K3).  Entries outside the pattern are
 * exactly 0.0 in every dense J the sweep kernel writes.                                        */
int ogb_jac_pattern(void* prob, uint32_t* lin_h, int cap);
int ogb_pack(void* prob, const double* J, int B, double* vals, void* stream);

/* ---- host-buffer entry point -----------------------------------------------------------
 * What the reference's SciPy-facing closures are to a host caller: decision vectors in HOST
 * memory in, c and the dense FD Jacobian in HOST memory out (reference optimize.py:711-715 +
 * scipy/optimize/_slsqp_py.py:353-367, for B instances at once).  A session owns the device
 * scratch, two CUDA streams, pinned staging and a pool of host threads for one problem on the
 * current device.  ogb_host_eval_fd is synchronous: when it returns, c_h [B, nrows] and J_h
 * are complete.  The batch is cut into chunks that flow through
 *     H2D p -> K1 -> K2a (sweep kernel, packed output) -> D2H packed values -> host threads
 * write the dense J_h (zeros with non-temporal stores + the packed non-zeros), so neither HBM nor
 * PCIe ever carries more than nnz doubles per instance and the copies overlap the kernels.
 * p_h / c_h / J_h may be pageable or pinned memory.                                        */
enum ogb_host_mode {
    OGB_HOST_J_DENSE = 0,      /* J_h [B, nvars, nrows] fully rewritten (zeros included)              */
    OGB_HOST_J_KEEP_ZEROS = 1, /* J_h [B, nvars, nrows] already holds this problem's zero background
                                  (e.g. from an earlier OGB_HOST_J_DENSE call into the same buffer):
                                  only the entries of the pattern are rewritten                       */
    OGB_HOST_J_PACKED = 2,     /* J_h [B, nnz] receives the packed values (pattern: ogb_jac_pattern)  */
    OGB_HOST_J_DMA = 3         /* J_h [B, nvars, nrows] written by one dense device->host copy (the
                                  transport without packing; fastest into pinned memory)              */
};

typedef struct ogb_host_stats {
    int64_t h2d_bytes, d2h_bytes;   /* bytes copied over PCIe by the last call                        */
    int32_t launches;               /* kernels launched by the last call                              */
    int32_t nnz, chunk, threads, nchunks, pad;
    double ms_total, ms_first_chunk;/* wall time of the last call; time until the first chunk landed  */
} ogb_host_stats;

void* ogb_host_session_create(void* prob, int max_batch, int chunk /*0 = default*/, int threads /*0 = all*/);
void  ogb_host_session_destroy(void* session);
int   ogb_host_eval_fd(void* session, const double* p_h, const double* lb_h, const double* ub_h,
                       double abs_step, int B, double* c_h, double* J_h, int mode);
int   ogb_host_session_stats(void* session, ogb_host_stats* out);
enum ogb_host_option {
    OGB_HOST_OPT_EXACT = 0     /* 1: the session's Jacobians are the exact ones (ogb_eval_exact) instead of SciPy's
                                  forward differences (ogb_eval_sparse); the packed modes only -- OGB_HOST_J_DMA
                                  stays the FD Jacobian                                                          */
};
int   ogb_host_session_set_option(void* session, int key, int value);

/* The same evaluation delivered straight into an SQP driver's own buffers -- what SciPy's
 * _eval_con_normals does per instance with `C[row:row+k, :] = jac(x)` (scipy/optimize/
 * _slsqp_py.py:599-615) and `g = sf.grad(x)` (:533): instance b's packed non-zeros are scattered into
 * the Fortran-ordered matrix C_h[b] (leading dimension ld >= mrows; entry (r, j) at C_h[b][j * ld + r])
 * for the rows r < mrows (mrows = meq + mineq: every constraint row), and the cost row (r = nrows - 1)
 * into the vector g_h[b][j] (g_h: B pointers, or NULL).  Only structural non-zeros are written: the matrices must
 * hold this problem's zero background (they do if they were zero-initialised and only ever written by
 * this call; SLSQP's core does not modify C).  c_h [B, nrows] as in ogb_host_eval_fd.              */
int   ogb_host_eval_fd_scatter(void* session, const double* p_h, const double* lb_h, const double* ub_h,
                               double abs_step, int B, double* c_h, double* const* C_h, int ld, int mrows,
                               double* const* g_h);

/* The host half of the transport by itself (no GPU involved): expand packed values [B, nnz] into
 * dense J_h [B, nM] (mode OGB_HOST_J_DENSE or OGB_HOST_J_KEEP_ZEROS) with `threads` threads.   */
int ogb_host_expand(const double* vals_h, const uint32_t* lin_h, int nnz, size_t nM, int B,
                    double* J_h, int mode, int threads);

/* ---- batched SLSQP on the device ("device-side batched QP") ------------------
 * What scipy.optimize.minimize(method="SLSQP") runs per instance for the reference's Problem.solve
 * (/root/reference/OpenGoddard/optimize.py:738-755; driver loop scipy/optimize/_slsqp_py.py:524-555), for B
 * instances at once: one thread block advances one instance through one reverse-communication step of
 * Kraft's SLSQP -- damped BFGS update of L D L', the QP as LSQ -> LSEI -> LSI -> LDP -> NNLS (augmented with the
 * slack variable when the linearisation is inconsistent), the L1 merit line search, the convergence tests
 * (csrc/ogb_sqp.h restates the published algorithm; SciPy's compiled core is not part of the reference tree).
 * The caller alternates   ogb_eval_sparse / ogb_eval_exact (c, vals at x)   and   ogb_sqp_step   until no
 * instance reports mode 1 or -1.  Opt-in: the default multi-start keeps SciPy's own core on the host.
 *
 * ogb_sqp_create: nvars, m constraints (the first meq are equalities; m >= 1, meq < nvars, at least one
 * inequality or finite bound), the packed Jacobian pattern BY VARIABLE (colptr_h [nvars + 1], prow_h [nnz], rows
 * in [0, m], row m = cost gradient: ogb_jac_pattern's ascending indices j * nrows + r, split), the bounds
 * (xl_h / xu_h [nvars], +-inf or NaN = none), SLSQP's acc (ftol) and iteration limit, and the largest batch.
 * ogb_sqp_start: (re)start B instances -- the next step treats c / vals as the values at the start points.
 * ogb_sqp_step: x [B, nvars] device, in / out; c [B, m + 1], vals [B, nnz] device, as written by the sweep
 * kernel at x; mode_h [B] host (may be NULL: no synchronisation): SLSQP's mode per instance after the step --
 * 1: evaluate c at x (line search), -1: evaluate c and the Jacobian at x, 0: converged, 2..9: SLSQP's exit modes.
 * ogb_sqp_scalars: [B, 24] doubles per instance to the host: f, f0, gs, h1, h2, h3, h4, t, t0, alpha, mode, iter,
 * reset, line, inconsistent, nfev, njev, (unused), SM cycles spent in the QP phases (6).                        */
void*  ogb_sqp_create(int nvars, int m, int meq, int nnz, const int32_t* colptr_h, const int32_t* prow_h,
                      const double* xl_h, const double* xu_h, double acc, int maxiter, int max_batch);
void   ogb_sqp_destroy(void* sqp);
size_t ogb_sqp_bytes(void* sqp);
int    ogb_sqp_start(void* sqp, int B, void* stream);
int    ogb_sqp_step(void* sqp, double* x, const double* c, const double* vals, int B, int32_t* mode_h, void* stream);
int    ogb_sqp_scalars(void* sqp, int B, double* sc_h, void* stream);
long long ogb_sqp_launches(void* sqp);

#ifdef __cplusplus
}
#endif
#endif /* OGB200_H */
