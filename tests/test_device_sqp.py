"""Device SQP ("device-side batched QP"): the oracle of this path is a numpy
restatement of Kraft's SLSQP (oracle/og_sqp.py, oracle/og_lsq.py).  It is pinned here against the installed SciPy --
`scipy.optimize.minimize(method="SLSQP")`, the one call the reference's Problem.solve makes
(/root/reference/OpenGoddard/optimize.py:738-755) -- before anything on the device is compared with it."""
import numpy as np
import pytest
from scipy import optimize

from oracle import og_lsq, og_sqp

INF = np.inf


def _problems():
    out = {}
    f = lambda x: (1 - x[0]) ** 2 + 100 * (x[1] - x[0] ** 2) ** 2
    gf = lambda x: np.array([-2 * (1 - x[0]) - 400 * x[0] * (x[1] - x[0] ** 2), 200 * (x[1] - x[0] ** 2)])
    out["rosenbrock_in_a_disk"] = (f, gf, lambda x: np.zeros(0), lambda x: np.zeros((0, 2)),
                                   lambda x: np.array([2 - x[0] ** 2 - x[1] ** 2]), lambda x: np.array([[-2 * x[0], -2 * x[1]]]),
                                   np.array([-1.2, 1.0]), np.full(2, -INF), np.full(2, INF), 100)
    f = lambda x: x[0] * x[3] * (x[0] + x[1] + x[2]) + x[2]
    gf = lambda x: np.array([x[3] * (2 * x[0] + x[1] + x[2]), x[0] * x[3], x[0] * x[3] + 1, x[0] * (x[0] + x[1] + x[2])])
    out["hs71"] = (f, gf, lambda x: np.array([x @ x - 40]), lambda x: 2 * x[None, :], lambda x: np.array([np.prod(x) - 25]),
                   lambda x: np.array([[x[1] * x[2] * x[3], x[0] * x[2] * x[3], x[0] * x[1] * x[3], x[0] * x[1] * x[2]]]),
                   np.array([1.0, 5, 5, 1]), np.ones(4), 5 * np.ones(4), 100)
    rng = np.random.default_rng(1)
    n = 12
    Q = rng.normal(size=(n, n))
    Q = Q @ Q.T + np.eye(n)
    q = rng.normal(size=n)
    Ae, Ai = rng.normal(size=(3, n)), rng.normal(size=(6, n))
    out["random_nlp"] = (lambda x: 0.5 * x @ Q @ x + q @ x + 0.1 * np.sum(x ** 4), lambda x: Q @ x + q + 0.4 * x ** 3,
                         lambda x: Ae @ x + 0.1 * np.sin(x[:3]) - 0.3,
                         lambda x: Ae + 0.1 * np.hstack([np.diag(np.cos(x[:3])), np.zeros((3, n - 3))]),
                         lambda x: Ai @ x + 1.0 - 0.05 * x[:6] ** 2,
                         lambda x: Ai - 0.1 * np.hstack([np.diag(x[:6]), np.zeros((6, n - 6))]),
                         rng.normal(size=n) * 0.3, -np.ones(n), np.ones(n), 100)
    # the first linearisation is inconsistent (gradient of x1^2 + x2^2 >= 1 vanishes at the start): augmented QP
    f = lambda x: (x[0] - 2) ** 2 + (x[1] + 1) ** 2 + x[2] ** 2
    gf = lambda x: np.array([2 * (x[0] - 2), 2 * (x[1] + 1), 2 * x[2]])
    inc = (f, gf, lambda x: np.array([x[0] - x[1] - 0.3 + x[2] ** 2]), lambda x: np.array([[1.0, -1.0, 2 * x[2]]]),
           lambda x: np.array([x[0] ** 2 + x[1] ** 2 - 1.0, -(x[0] ** 2 + x[1] ** 2) + 0.25 + 3 * x[2] ** 2]),
           lambda x: np.array([[2 * x[0], 2 * x[1], 0.0], [-2 * x[0], -2 * x[1], 6 * x[2]]]),
           np.array([0.0, 0.0, 0.1]), -3 * np.ones(3), 3 * np.ones(3))
    out["inconsistent_linearisation"] = inc + (100,)
    out["iteration_limit"] = inc + (3,)
    return out


PROBLEMS = _problems()


def callables(name):
    f, gf, ceq, jeq, cin, jin, x0, lb, ub, maxiter = PROBLEMS[name]
    n = len(x0)
    evalf = lambda x: (f(x), np.concatenate([ceq(x), cin(x)]))
    evalg = lambda x: (gf(x), np.vstack([jeq(x).reshape(-1, n), jin(x).reshape(-1, n)]))
    return evalf, evalg, len(ceq(x0))


@pytest.mark.parametrize("name", sorted(PROBLEMS))
def test_restated_slsqp_reproduces_scipy_iterate_by_iterate(name):
    f, gf, ceq, jeq, cin, jin, x0, lb, ub, maxiter = PROBLEMS[name]
    evalf, evalg, meq = callables(name)
    cons = ([{"type": "eq", "fun": ceq, "jac": jeq}] if meq else []) + [{"type": "ineq", "fun": cin, "jac": jin}]
    bounds = list(zip(np.where(np.isfinite(lb), lb, None), np.where(np.isfinite(ub), ub, None)))
    trial = []                                            # SciPy's callback sees the first trial point x_k + s_k of every iteration
    ref = optimize.minimize(f, x0, jac=gf, method="SLSQP", bounds=bounds, constraints=cons,
                            options={"ftol": 1e-9, "maxiter": maxiter}, callback=lambda x: trial.append(x.copy()))
    trace = []
    out = og_sqp.slsqp_numpy(evalf, evalg, x0, lb, ub, meq, 1e-9, maxiter, trace=trace)
    assert (out["status"], out["nit"], out["nfev"], out["njev"]) == (ref.status, ref.nit, ref.nfev, ref.njev)
    assert np.abs(out["x"] - ref.x).max() <= 1e-9 * max(1.0, np.abs(ref.x).max())
    assert abs(out["fun"] - ref.fun) <= 1e-10 * max(1.0, abs(ref.fun))
    for (it, x, s, r, h4), xt in zip(trace, trial):
        assert np.abs(x + s - xt).max() <= 1e-7 * max(1.0, np.abs(xt).max()), it
    if name == "inconsistent_linearisation":
        assert trace[0][4] < 1.0                          # (the slack variable was used)


def _kkt_ok(L, Dg, g, A, c, meq, lo, hi, x, y, tol=1e-8):
    Bm = L @ np.diag(Dg) @ L.T
    s = A @ x + c
    res = Bm @ x + g - A.T @ y                            # = multipliers of the bounds
    at_lo, at_hi = x <= lo + 1e-9, x >= hi - 1e-9
    free = ~at_lo & ~at_hi
    return (np.abs(res[free]).max(initial=0.0) < tol and np.abs(s[:meq]).max(initial=0.0) < tol
            and s[meq:].min(initial=0.0) > -tol and y[meq:].min(initial=0.0) > -tol
            and np.abs(y[meq:] * s[meq:]).max(initial=0.0) < tol
            and (res[at_lo & ~at_hi] >= -tol).all() and (res[at_hi & ~at_lo] <= tol).all())


def random_qp(rng):
    n = int(rng.integers(2, 30))
    meq = int(rng.integers(0, max(1, n - 2)))
    mi = int(rng.integers(1, 40))
    m = meq + mi
    L = np.tril(rng.normal(size=(n, n)) * 0.3, -1) + np.eye(n)
    Dg = rng.uniform(0.1, 3.0, n)
    g = rng.normal(size=n)
    A = rng.normal(size=(m, n))
    xf = rng.normal(size=n)
    c = -A @ xf
    c[meq:] += rng.uniform(0.05, 1.0, mi)                  # xf is strictly feasible
    lo = np.where(rng.random(n) < 0.5, xf - rng.uniform(0.05, 1, n), -INF)
    hi = np.where(rng.random(n) < 0.5, xf + rng.uniform(0.05, 1, n), INF)
    return L, Dg, g, A, c, meq, lo, hi


def test_lsq_solves_random_qps_to_their_kkt_conditions():
    rng = np.random.default_rng(0)
    for trial in range(150):
        L, Dg, g, A, c, meq, lo, hi = random_qp(rng)
        x, y, mode = og_lsq.lsq(L, Dg, g, A, c, meq, lo, hi)
        assert mode == 1, trial
        assert _kkt_ok(L, Dg, g, A, c, meq, lo, hi, x, y), trial


def test_lsq_step_is_scipys_step_on_the_first_iteration():
    """B = I on SLSQP's first iteration: SciPy's low-level step solves the very QP og_lsq.lsq is given."""
    from opengoddard_b200 import sqp
    slsqp, ilp64 = sqp._low_level()
    rng = np.random.default_rng(3)
    for trial in range(40):
        L, Dg, g, A, c, meq, lo, hi = random_qp(rng)
        n, m = len(g), len(c)
        x, y, mode = og_lsq.lsq(np.eye(n), np.ones(n), g, A, c, meq, lo, hi)
        it = sqp._Instance(np.zeros(n), n, m, meq, 1e-6, 25, np.int64 if ilp64 else np.int32)
        it.g[:] = g
        it.C[:m] = A
        it.d[:m] = c
        slsqp(it.state, 0.0, it.g, it.C, it.d, it.x, it.mult, np.where(np.isfinite(lo), lo, np.nan),
              np.where(np.isfinite(hi), hi, np.nan), it.buffer, it.indices)
        assert it.state["mode"] == 1 and it.state["h4"] == 1.0 and mode == 1
        assert np.abs(it.x - x).max() <= 1e-9 * max(1.0, np.abs(x).max()), trial
