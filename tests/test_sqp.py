"""Batched SQP driver (opengoddard_b200/sqp.py): host logic against SciPy itself, with the oracle
standing in for the device evaluator (the GPU variant is in test_gpu_parity.py)."""
import numpy as np
import pytest
from scipy import optimize

from opengoddard_b200 import sqp, workloads
from oracle import og_numpy


class OracleEvaluator:
    def __init__(self, wl):
        self.wl = wl
        self.lb, self.ub = og_numpy.bounds_arrays(wl.prob)
        self.calls = []

    def eval(self, X):
        self.calls.append(("f", len(X)))
        return np.stack([og_numpy.eval_c(self.wl.prob, self.wl.obj, x) for x in X])

    def eval_fd(self, X):
        self.calls.append(("g", len(X)))
        cs, Js = [], []
        for x in X:
            c, J = og_numpy.eval_fd(self.wl.prob, self.wl.obj, x, self.lb, self.ub)
            cs.append(c)
            Js.append(np.ascontiguousarray(J.T))
        return np.stack(cs), np.stack(Js)


def scipy_reference(wl, ev, x0, meq, mineq, ftol, maxiter):
    """scipy.optimize.minimize on the same callables (contiguous gradients)."""
    M = meq + mineq
    c_at = lambda x: og_numpy.eval_c(wl.prob, wl.obj, x)
    j_at = lambda x: og_numpy.eval_fd(wl.prob, wl.obj, x, ev.lb, ev.ub)[1]
    cons = ({"type": "eq", "fun": lambda x: c_at(x)[:meq], "jac": lambda x: j_at(x)[:meq]},
            {"type": "ineq", "fun": lambda x: c_at(x)[meq:M], "jac": lambda x: j_at(x)[meq:M]})
    return optimize.minimize(lambda x: c_at(x)[M], x0, jac=lambda x: np.ascontiguousarray(j_at(x)[M]),
                             bounds=wl.prob.bounds, constraints=cons, method="SLSQP",
                             options={"maxiter": maxiter, "ftol": ftol})


def test_batch_of_one_reproduces_scipy_minimize():
    wl = workloads.build("cfg1_brachistochrone20", og_numpy)
    ev = OracleEvaluator(wl)
    meq, mineq = 64, 41
    x0 = wl.prob.p.copy()
    res = sqp.slsqp_batch(ev, x0[None], ev.lb, ev.ub, meq, mineq, ftol=1e-6, maxiter=12)
    ref = scipy_reference(wl, ev, x0, meq, mineq, 1e-6, 12)
    assert res["status"][0] == ref.status and res["nit"][0] == ref.nit
    assert np.array_equal(res["x"][0], ref.x)
    assert res["fun"][0] == ref.fun


def test_instances_are_independent_and_batched():
    wl = workloads.build("cfg1_brachistochrone20", og_numpy)
    ev = OracleEvaluator(wl)
    meq, mineq = 64, 41
    P = np.vstack([wl.prob.p[None], workloads.make_batch(wl, 2)])
    res = sqp.slsqp_batch(ev, P, ev.lb, ev.ub, meq, mineq, ftol=1e-6, maxiter=6)
    assert ev.calls[0] == ("g", 3)                          # one batched call serves all instances
    assert max(k for _, k in ev.calls) == 3
    for b in range(3):
        one = sqp.slsqp_batch(OracleEvaluator(wl), P[b][None], ev.lb, ev.ub, meq, mineq, ftol=1e-6, maxiter=6)
        assert np.array_equal(one["x"][0], res["x"][b]) and one["status"][0] == res["status"][b]
    two = sqp.slsqp_batch(OracleEvaluator(wl), P, ev.lb, ev.ub, meq, mineq, ftol=1e-6, maxiter=6, threads=2)
    assert np.array_equal(two["x"], res["x"])


def test_converges_like_the_reference_outer_loop():
    """Restarting unfinished instances (reference optimize.py:738-755) reaches t_f = sqrt(pi)."""
    wl = workloads.build("cfg1_brachistochrone20", og_numpy)
    ev = OracleEvaluator(wl)
    meq, mineq = 64, 41
    X = wl.prob.p[None].copy()
    grad = lambda x: wl.prob.eval_cost_derivative(x, wl.obj)
    for _ in range(6):
        res = sqp.slsqp_batch(ev, X, ev.lb, ev.ub, meq, mineq, ftol=1e-6, maxiter=25, cost_grad=grad)
        X = res["x"]
        if res["status"][0] == 0:
            break
    assert res["status"][0] == 0
    assert abs(X[0, -1] - 1.7724608832526498) < 1e-4


def test_scipy_slsqp_needs_contiguous_gradient():
    """Documented SciPy 1.18 behaviour the facade guards against: the low-level step reads a
    strided gradient array as if it were contiguous."""
    slsqp, _ = sqp._low_level()
    wl = workloads.build("cfg1_brachistochrone20", og_numpy)
    lb, ub = og_numpy.bounds_arrays(wl.prob)
    n, meq, m = 81, 64, 105
    x0 = np.clip(wl.prob.p, lb, ub)
    c, J = og_numpy.eval_fd(wl.prob, wl.obj, x0, lb, ub)

    def one_step(g):
        it = sqp._Instance(x0, n, m, meq, 1e-6, 2, np.int32)
        it.C[:m] = J[:m]
        it.d[:m] = c[:m]
        xl = np.where(np.isfinite(lb), lb, np.nan)
        xu = np.where(np.isfinite(ub), ub, np.nan)
        slsqp(it.state, float(c[m]), g, it.C, it.d, it.x, it.mult, xl, xu, it.buffer, it.indices)
        return it.x

    g = np.ascontiguousarray(J[m])
    strided = np.ascontiguousarray(J.T)[:, m]
    assert np.array_equal(g, strided) and not strided.flags["C_CONTIGUOUS"]
    good, bad = one_step(g), one_step(strided)
    if np.array_equal(good, bad):
        pytest.skip("this SciPy handles strided gradients")
    assert not np.array_equal(good, bad)


def test_process_parallel_stepping_is_bitwise_identical():
    """processes=2: the SLSQP states live in worker subprocesses (shared-memory exchange); same result
    as one process with a single BLAS thread, same batched evaluator calls, with and without a user
    cost gradient."""
    wl = workloads.build("cfg1_brachistochrone20", og_numpy)
    meq, mineq = 64, 41
    P = np.vstack([wl.prob.p[None], workloads.make_batch(wl, 4)])
    grad = lambda x: wl.prob.eval_cost_derivative(x, wl.obj)
    for cg in (None, grad):
        ev1, ev2 = OracleEvaluator(wl), OracleEvaluator(wl)
        seen = []
        from threadpoolctl import threadpool_limits
        with threadpool_limits(1):                  # the workers run SLSQP's LAPACK with one BLAS thread
            one = sqp.slsqp_batch(ev1, P, ev1.lb, ev1.ub, meq, mineq, ftol=1e-6, maxiter=5, cost_grad=cg)
        two = sqp.slsqp_batch(ev2, P, ev2.lb, ev2.ub, meq, mineq, ftol=1e-6, maxiter=5, cost_grad=cg, processes=2,
                              callback=lambda b, x, f: seen.append(b))
        for key in ("x", "fun", "status", "nit", "nfev", "njev"):
            assert np.array_equal(one[key], two[key]), key
        assert one["message"] == two["message"] and ev1.calls == ev2.calls
        assert set(seen) == set(range(len(P)))


def test_worker_pool_is_reusable_across_solves():
    """One WorkerPool serves several slsqp_batch calls (different batch sizes), like solve_batch's passes."""
    from threadpoolctl import threadpool_limits
    wl = workloads.build("cfg1_brachistochrone20", og_numpy)
    meq, mineq = 64, 41
    P = workloads.make_batch(wl, 5)
    with sqp.WorkerPool(3) as pool:
        for rows in (P, P[:2], P[1:2]):
            ev1, ev2 = OracleEvaluator(wl), OracleEvaluator(wl)
            with threadpool_limits(1):
                one = sqp.slsqp_batch(ev1, rows, ev1.lb, ev1.ub, meq, mineq, ftol=1e-6, maxiter=3)
            two = sqp.slsqp_batch(ev2, rows, ev2.lb, ev2.ub, meq, mineq, ftol=1e-6, maxiter=3, processes=pool)
            assert np.array_equal(one["x"], two["x"]) and np.array_equal(one["status"], two["status"])
        assert all(pr.poll() is None for pr in pool.procs)       # still alive between calls
    assert pool.procs == []


class ScatterEvaluator(OracleEvaluator):
    """Like the device evaluator (engine._HostEvaluator.eval_fd_scatter): delivers the Jacobians straight
    into the SQP driver's own per-instance buffers, addressed by raw pointers -- C_ptrs[k] a Fortran-ordered
    (ld, n) matrix receiving the rows < mrows, G_ptrs[k] the gradient vector receiving the cost row."""

    def __init__(self, wl):
        super().__init__(wl)
        self.direct = 0

    def eval_fd_scatter(self, X, C_ptrs, ld, mrows, G_ptrs=None):
        import ctypes
        c, J = super().eval_fd(X)                      # J (k, n, m + 1)
        n = J.shape[1]
        for k in range(len(X)):
            Cv = np.ctypeslib.as_array(ctypes.cast(int(C_ptrs[k]), ctypes.POINTER(ctypes.c_double)), shape=(n, ld))
            Cv[:, :mrows] = J[k, :, :mrows]
            if G_ptrs is not None:
                gv = np.ctypeslib.as_array(ctypes.cast(int(G_ptrs[k]), ctypes.POINTER(ctypes.c_double)), shape=(n,))
                gv[:] = J[k, :, -1]
        self.direct += 1
        return c


def test_evaluator_writes_jacobians_straight_into_shared_memory():
    """An evaluator that offers eval_fd_scatter writes the Jacobians into the SLSQP states' own buffers (the
    workers' shared memory, or the local instances' C / g arrays): same iterates as the copying path."""
    from threadpoolctl import threadpool_limits
    wl = workloads.build("cfg1_brachistochrone20", og_numpy)
    meq, mineq = 64, 41
    P = workloads.make_batch(wl, 4)
    ev1, ev2 = OracleEvaluator(wl), ScatterEvaluator(wl)
    with threadpool_limits(1):
        one = sqp.slsqp_batch(ev1, P, ev1.lb, ev1.ub, meq, mineq, ftol=1e-6, maxiter=5)
    two = sqp.slsqp_batch(ev2, P, ev2.lb, ev2.ub, meq, mineq, ftol=1e-6, maxiter=5, processes=2)
    assert ev2.direct >= 1
    for key in ("x", "fun", "status", "nit"):
        assert np.array_equal(one[key], two[key]), key
    ev4 = ScatterEvaluator(wl)
    with threadpool_limits(1):
        four = sqp.slsqp_batch(ev4, P, ev1.lb, ev1.ub, meq, mineq, ftol=1e-6, maxiter=5)        # local stepper
    assert ev4.direct >= 1 and np.array_equal(four["x"], one["x"])
    # a user cost gradient takes precedence over the evaluator's cost row
    grad = lambda x: og_numpy.eval_fd(wl.prob, wl.obj, x, ev1.lb, ev1.ub)[1][-1]
    with threadpool_limits(1):
        five = sqp.slsqp_batch(ScatterEvaluator(wl), P[:2], ev1.lb, ev1.ub, meq, mineq, ftol=1e-6, maxiter=3, cost_grad=grad)
        six = sqp.slsqp_batch(OracleEvaluator(wl), P[:2], ev1.lb, ev1.ub, meq, mineq, ftol=1e-6, maxiter=3, cost_grad=grad)
    assert np.array_equal(five["x"], six["x"])
