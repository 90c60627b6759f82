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


# ---------------------------------------------------------------- the device code, run by one serial thread
def dense_pattern(n, M):
    return np.arange(n + 1, dtype=np.int32) * M, np.tile(np.arange(M, dtype=np.int32), n)


def pack_dense(A, g):
    """(m, n) constraint normals + (n,) cost gradient -> packed values in the dense pattern (by variable)."""
    return np.ascontiguousarray(np.vstack([A, g[None, :]]).T).ravel()


def test_device_lsq_matches_the_restatement_on_random_qps():
    from tests.emu.emu import EmuSqp
    rng = np.random.default_rng(5)
    for trial in range(60):
        L, Dg, g, A, c, meq, lo, hi = random_qp(rng)
        n, m = len(g), len(c)
        x0 = rng.normal(size=n)
        colptr, prow = dense_pattern(n, m + 1)
        emu = EmuSqp(n, m, meq, colptr, prow, lo + x0, hi + x0, 1e-6, 10, 1)
        emu.field(0, "lt", n * n)[:] = np.ascontiguousarray(L.T).ravel()
        emu.field(0, "dg", n)[:] = Dg
        emu.field(0, "g", n)[:] = g
        cc = np.concatenate([c, [0.0]])
        mode = emu.lsq(0, x0.copy(), cc, pack_dense(A, g))
        x, y, mode_ref = og_lsq.lsq(L, Dg, g, A, c, meq, lo, hi)
        assert mode == mode_ref == 1, trial
        assert np.abs(emu.field(0, "s", n) - x).max() <= 1e-9 * max(1.0, np.abs(x).max()), trial
        assert np.abs(emu.field(0, "r", m) - y).max() <= 1e-8 * max(1.0, np.abs(y).max()), trial


def drive(emu, evalf, evalg, x0, lb, ub, rounds=2000):
    """The reverse-communication loop around the device step (what DeviceSqp does with the GPU evaluator)."""
    n, m = emu.n, emu.m
    X = np.clip(np.asarray(x0, dtype=float), lb, ub)[None, :].copy()
    for _ in range(rounds):
        xc = np.clip(X[0], lb, ub)
        f, c = evalf(xc)
        g, A = evalg(xc)
        emu.step(X, np.concatenate([c, [f]])[None, :].copy(), pack_dense(A, g)[None, :].copy())
        sc = emu.scalars(0)
        if abs(sc["mode"]) != 1:
            break
    f, _ = evalf(np.clip(X[0], lb, ub))
    return {"x": X[0], "fun": f, "status": int(sc["mode"]), "nit": int(sc["iter"]), "nfev": int(sc["nfev"]),
            "njev": int(sc["njev"])}


@pytest.mark.parametrize("name", sorted(PROBLEMS))
def test_device_sqp_matches_the_restatement(name):
    from tests.emu.emu import EmuSqp
    f, gf, ceq, jeq, cin, jin, x0, lb, ub, maxiter = PROBLEMS[name]
    evalf, evalg, meq = callables(name)
    n = len(x0)
    m = len(evalf(x0)[1])
    colptr, prow = dense_pattern(n, m + 1)
    emu = EmuSqp(n, m, meq, colptr, prow, lb, ub, 1e-9, maxiter, 1)
    out = drive(emu, evalf, evalg, x0, lb, ub)
    ref = og_sqp.slsqp_numpy(evalf, evalg, x0, lb, ub, meq, 1e-9, maxiter)
    assert (out["status"], out["nit"], out["nfev"], out["njev"]) == (ref["status"], ref["nit"], ref["nfev"], ref["njev"])
    assert np.abs(out["x"] - ref["x"]).max() <= 1e-8 * max(1.0, np.abs(ref["x"]).max())


def test_solve_batch_rejects_an_unknown_qp_before_touching_the_gpu(api):
    from opengoddard_b200 import workloads
    wl = workloads.build("cfg1_brachistochrone20", api)
    with pytest.raises(ValueError):
        wl.prob.solve_batch(np.asarray(wl.prob.p)[None], wl.obj, qp="cvx")


# ---------------------------------------------------------------- the CUDA kernel (thread block per instance)
def _kernel(n, m, meq, colptr, prow, lb, ub, ftol, maxiter, B):
    import torch
    from opengoddard_b200 import capi, engine
    dev = torch.device("cuda", torch.cuda.current_device())
    return engine.SqpKernel(capi.ogb(), torch, dev, n, m, meq, colptr, prow, lb, ub, ftol, maxiter, B)


@pytest.mark.gpu
def test_kernel_first_step_solves_a_batch_of_different_qps():
    """One launch, 96 instances with different data: the first SLSQP step (B = I) must leave x0 + (QP solution) in
    x for every instance -- checked against the numpy restatement, which the tests above pin to SciPy."""
    import torch
    rng = np.random.default_rng(11)
    n, meq, mi = 24, 7, 30
    m = meq + mi
    colptr, prow = dense_pattern(n, m + 1)
    B = 96
    lo_mask, hi_mask = rng.random(n) < 0.5, rng.random(n) < 0.5
    lb = np.where(lo_mask, -1.5, -INF)
    ub = np.where(hi_mask, 1.5, INF)
    X0 = rng.uniform(-1.0, 1.0, (B, n))
    cs, vs, ref = [], [], []
    for b in range(B):
        g = rng.normal(size=n)
        A = rng.normal(size=(m, n))
        xf = np.clip(X0[b] + rng.normal(size=n) * 0.3, -1.4, 1.4)         # a strictly feasible point of the linearisation
        c = -A @ (xf - X0[b])
        c[meq:] += rng.uniform(0.05, 1.0, mi)
        x, y, mode = og_lsq.lsq(np.eye(n), np.ones(n), g, A, c, meq, lb - X0[b], ub - X0[b])
        assert mode == 1
        ref.append(X0[b] + x)
        cs.append(np.concatenate([c, [0.0]]))
        vs.append(pack_dense(A, g))
    k = _kernel(n, m, meq, colptr, prow, lb, ub, 1e-6, 5, B)
    X = torch.from_numpy(X0.copy()).cuda()
    mode = np.zeros(B, dtype=np.int32)
    k.start(B)
    k.step(X, torch.from_numpy(np.stack(cs)).cuda(), torch.from_numpy(np.stack(vs)).cuda(), mode)
    assert (mode == 1).all()
    err = np.abs(X.cpu().numpy() - np.stack(ref)).max()
    assert err <= 1e-9, err
    sc = k.scalars(B)
    assert (sc["iter"] == 1).all() and (sc["h4"] == 1.0).all()
    k.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(PROBLEMS))
def test_kernel_runs_whole_solves_like_the_restatement(name):
    """Whole SLSQP runs on the GPU (host-evaluated f, c, g, A uploaded every round), several copies of the instance
    in one batch: status / nit / nfev / njev and x as the numpy restatement (and hence SciPy) produce them."""
    import torch
    f, gf, ceq, jeq, cin, jin, x0, lb, ub, maxiter = PROBLEMS[name]
    evalf, evalg, meq = callables(name)
    n = len(x0)
    m = len(evalf(x0)[1])
    colptr, prow = dense_pattern(n, m + 1)
    B = 3
    starts = [x0, x0 * 1.02 + 0.01, x0 * 1.04 + 0.02]
    refs = [og_sqp.slsqp_numpy(evalf, evalg, s, lb, ub, meq, 1e-9, maxiter) for s in starts]
    k = _kernel(n, m, meq, colptr, prow, lb, ub, 1e-9, maxiter, B)
    X = torch.from_numpy(np.clip(np.stack(starts), lb, ub)).cuda()
    mode = np.zeros(B, dtype=np.int32)
    k.start(B)
    for _ in range(3000):
        Xh = np.clip(X.cpu().numpy(), lb, ub)
        cs = np.stack([np.concatenate([evalf(x)[1], [evalf(x)[0]]]) for x in Xh])
        vs = np.stack([pack_dense(evalg(x)[1], evalg(x)[0]) for x in Xh])
        k.step(X, torch.from_numpy(cs).cuda(), torch.from_numpy(vs).cuda(), mode)
        if not (np.abs(mode) == 1).any():
            break
    sc = k.scalars(B)
    Xh = X.cpu().numpy()
    for b, ref in enumerate(refs):
        got = (int(sc["mode"][b]), int(sc["iter"][b]), int(sc["nfev"][b]), int(sc["njev"][b]))
        assert got == (ref["status"], ref["nit"], ref["nfev"], ref["njev"]), (b, got)
        assert np.abs(Xh[b] - ref["x"]).max() <= 1e-7 * max(1.0, np.abs(ref["x"]).max()), b
    k.close()


@pytest.mark.gpu
def test_device_sqp_on_collocation_problems(api):
    """Problem.solve_batch(qp="device") on Goddard-50 multi-starts: every instance ends in SLSQP's success or
    iteration-limit exit with the constraints satisfied and a cost at least as good as the SciPy-core path reaches
    from the same starts (the two are the same algorithm; on these ill-conditioned subproblems SciPy >= 1.16's
    compiled least-squares core resolves near-dependent constraints differently, see oracle/og_sqp.py)."""
    from opengoddard_b200 import workloads
    wl = workloads.build("cfg2_goddard50", api)
    P0 = workloads.make_batch(wl, 12)
    dev = wl.prob.solve_batch(P0, wl.obj, maxiter=25, max_outer=2, qp="device")
    ref = wl.prob.solve_batch(P0, wl.obj, maxiter=25, max_outer=2)
    eng = wl.prob.compile(wl.obj)
    c = eng.eval(np.clip(dev["x"], *wl.prob.bounds_arrays())).cpu().numpy()
    viol = np.array([og_sqp.violation(ci[:-1], eng.meq) for ci in c])
    assert np.isin(dev["status"], (0, 9)).all(), dev["status"]
    assert (viol < 5e-3).all(), viol                          # (status 9 instances are still iterating)
    assert (dev["fun"] <= ref["fun"] + 2e-3).all(), (dev["fun"], ref["fun"])


@pytest.mark.gpu
def test_device_sqp_first_iterations_match_the_serial_emulation(api):
    """The collocation path end to end (packed sweep -> SLSQP step kernel) against the same code run by one serial
    host thread on the host-compiled evaluator, two iterations from the reference's own initial guess."""
    from opengoddard_b200 import tape, workloads
    from oracle import og_numpy
    from tests.emu.emu import EmuProblem, EmuSqp
    wl = workloads.build("cfg2_goddard50", api)
    wo = workloads.build("cfg2_goddard50", og_numpy)
    lb, ub = og_numpy.bounds_arrays(wo.prob)
    eng = wl.prob.compile(wl.obj)
    x0 = np.clip(np.asarray(wo.prob.p, dtype=float), lb, ub)
    with eng.device_sqp(2, 1e-6, 2) as dq:
        dev = dq.solve(np.stack([x0, x0]))
    lin = eng.jac_pattern().astype(np.int64)
    n, M = eng.nvars, eng.nrows
    colptr = np.searchsorted(lin, np.arange(n + 1) * M).astype(np.int32)
    prow = (lin % M).astype(np.int32)
    ep = EmuProblem(tape.build_ir(wl.prob, wl.obj), lb, ub)
    emu = EmuSqp(n, M - 1, eng.meq, colptr, prow, lb, ub, 1e-6, 2, 1)
    X = x0[None].copy()
    for _ in range(200):
        cc, JJ = ep.eval_fd(X)
        emu.step(X, cc, np.ascontiguousarray(JJ.reshape(1, -1)[:, lin]))
        if abs(emu.scalars(0)["mode"]) != 1:
            break
    sc = emu.scalars(0)
    assert (dev["status"] == int(sc["mode"])).all() and (dev["nit"] == int(sc["iter"])).all()
    assert (dev["nfev"] == int(sc["nfev"])).all()
    assert np.array_equal(dev["x"][0], dev["x"][1])
    assert np.abs(dev["x"][0] - X[0]).max() <= 1e-5 * np.abs(X[0]).max()
