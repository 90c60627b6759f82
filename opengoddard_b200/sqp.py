"""Batched SQP driver: B independent SLSQP runs advanced in lock step.

The reference solves one instance at a time: `scipy.optimize.minimize(method='SLSQP')`
drives Kraft's SLSQP through a reverse-communication loop and, on every "gradient
required" step, spends 91 % of its time finite-differencing the Python callbacks
(/root/reference/OpenGoddard/optimize.py:738-755; scipy/optimize/_slsqp_py.py:524-555).
Here the same loop is written once for a whole batch: every instance keeps its own SLSQP
state and is stepped with SciPy's low-level `_slsqplib.slsqp` (one QP / line-search step per
call, on the host), and the function / Jacobian evaluations that the instances request are
gathered and served by ONE batched device call (`ogb_eval` / `ogb_eval_fd`).  The arithmetic
per instance is exactly SciPy's (same C core, same state machine, same `acc`/`maxiter`
semantics, same clipping of x into the bounds), so a batch of one reproduces
`minimize(method='SLSQP')` driven by the device callables.

The evaluator only needs two methods taking/returning host arrays:
    eval(X (k, n))     -> c (k, m + 1)                     [c_eq ; c_ineq ; cost]
    eval_fd(X (k, n))  -> c (k, m + 1), J (k, n, m + 1)    J[i, j, :] = column j
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np

EXIT_MODES = {-1: "Gradient evaluation required (g & a)", 0: "Optimization terminated successfully",
              1: "Function evaluation required (f & c)", 2: "More equality constraints than independent variables",
              3: "More than 3*n iterations in LSQ subproblem", 4: "Inequality constraints incompatible",
              5: "Singular matrix E in LSQ subproblem", 6: "Singular matrix C in LSQ subproblem",
              7: "Rank-deficient equality constraint subproblem HFTI",
              8: "Positive directional derivative for linesearch", 9: "Iteration limit reached"}


def _low_level():
    try:
        from scipy.optimize._slsqplib import slsqp
        from scipy.linalg.lapack import HAS_ILP64
    except Exception as e:                       # pragma: no cover
        raise NotImplementedError("the batched SQP driver needs SciPy's low-level SLSQP step "
                                  "(scipy >= 1.16): %r" % (e,))
    return slsqp, HAS_ILP64


class _Instance:
    """SLSQP state of one problem instance (mirrors scipy/optimize/_slsqp_py.py:453-520)."""

    def __init__(self, x0, n, m, meq, acc, maxiter, int_dtype):
        mieq = m - meq
        self.x = np.array(x0, dtype=np.float64)
        self.state = {"acc": acc, "alpha": 0.0, "f0": 0.0, "gs": 0.0, "h1": 0.0, "h2": 0.0, "h3": 0.0,
                      "h4": 0.0, "t": 0.0, "t0": 0.0, "tol": 10.0 * acc, "exact": 0, "inconsistent": 0,
                      "reset": 0, "iter": 0, "itermax": int(maxiter), "line": 0, "m": m, "meq": meq,
                      "mode": 0, "n": n}
        self.indices = np.zeros([max(m + 2 * n + 2, 1)], dtype=int_dtype)
        size = (n * (n + 1) // 2 + 3 * m * n - (m + 5 * n + 7) * meq + 9 * m + 8 * n * n + 35 * n
                + meq * meq + 28)
        if mieq == 0:
            size += 2 * n * (n + 1)
        self.buffer = np.zeros(max(size, 1), dtype=np.float64)
        self.mult = np.zeros([max(1, m + 2 * n + 2)], dtype=np.float64)
        self.C = np.zeros([max(1, m), n], dtype=np.float64, order="F")
        self.d = np.zeros([max(1, m)], dtype=np.float64)
        self.g = np.zeros(n, dtype=np.float64)
        self.fx = 0.0
        self.nfev = self.njev = 0

    @property
    def mode(self):
        return self.state["mode"]


def slsqp_batch(evaluator, X0, lb, ub, meq, mineq, ftol=1e-6, maxiter=25, cost_grad=None,
                threads=1, callback=None):
    """Run SLSQP on every row of X0 in lock step.

    evaluator : object with eval(X) and eval_fd(X) (see module docstring)
    lb, ub    : (n,) bounds with +-inf for "none" (SciPy clips x0 into them first)
    cost_grad : optional callable x -> (n,) user gradient of the cost (reference
                `cost_derivative`, optimize.py:730-733); default = the FD row of J
    threads   : host threads stepping the per-instance QP cores
    Returns dict(x (B, n), fun (B,), status (B,), nit (B,), nfev, njev, message list).
    """
    slsqp, ilp64 = _low_level()
    X0 = np.atleast_2d(np.asarray(X0, dtype=np.float64))
    B, n = X0.shape
    m = int(meq + mineq)
    lb = np.asarray(lb, dtype=np.float64)
    ub = np.asarray(ub, dtype=np.float64)
    X0 = np.clip(X0, lb, ub)                                 # _slsqp_py.py:322
    xl = np.where(np.isfinite(lb), lb, np.nan)               # the C core wants NaN for "no bound"
    xu = np.where(np.isfinite(ub), ub, np.nan)
    inst = [_Instance(X0[b], n, m, int(meq), float(ftol), maxiter, np.int64 if ilp64 else np.int32)
            for b in range(B)]

    def put_values(ids, c):
        for k, b in enumerate(ids):
            it = inst[b]
            it.fx = float(c[k, m])
            it.d[:m] = c[k, :m]
            it.nfev += 1

    def put_normals(ids, J):
        for k, b in enumerate(ids):
            it = inst[b]
            it.C[:m, :] = J[k, :, :m].T
            it.g[:] = cost_grad(it.x) if cost_grad is not None else J[k, :, m]
            it.njev += 1

    # mode 0 on entry: objective, constraints and gradients at the start point
    ids = list(range(B))
    c, J = evaluator.eval_fd(np.stack([inst[b].x for b in ids]))
    put_values(ids, c)
    put_normals(ids, J)

    def step(b):
        it = inst[b]
        slsqp(it.state, it.fx, it.g, it.C, it.d, it.x, it.mult, xl, xu, it.buffer, it.indices)
        return b

    active = list(range(B))
    pool = ThreadPoolExecutor(threads) if threads > 1 else None
    iters_prev = [0] * B
    try:
        while active:
            if pool is not None:
                list(pool.map(step, active))
            else:
                for b in active:
                    step(b)
            need_f = [b for b in active if inst[b].mode == 1]
            need_g = [b for b in active if inst[b].mode == -1]
            if need_f:
                # SciPy clips x for the objective only (_clip_x_for_func); x stays inside the
                # bounds in exact arithmetic, so the clip matters for 1-2 ulp excursions
                Xf = np.stack([np.clip(inst[b].x, lb, ub) for b in need_f])
                put_values(need_f, evaluator.eval(Xf))
            if need_g:
                Xg = np.stack([inst[b].x for b in need_g])
                _, Jg = evaluator.eval_fd(Xg)
                put_normals(need_g, Jg)
            if callback is not None:
                for b in active:
                    if inst[b].state["iter"] > iters_prev[b]:
                        callback(b, inst[b].x, inst[b].fx)
                    iters_prev[b] = inst[b].state["iter"]
            active = [b for b in active if abs(inst[b].mode) == 1]
    finally:
        if pool is not None:
            pool.shutdown()
    return {"x": np.stack([it.x for it in inst]), "fun": np.array([it.fx for it in inst]),
            "status": np.array([it.mode for it in inst]), "nit": np.array([it.state["iter"] for it in inst]),
            "nfev": np.array([it.nfev for it in inst]), "njev": np.array([it.njev for it in inst]),
            "message": [EXIT_MODES.get(it.mode, "?") for it in inst]}
