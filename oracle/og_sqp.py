"""TEST INFRASTRUCTURE (oracle): numpy restatement of Kraft's SLSQP outer iteration (`slsqpb`) on top of the
least-squares core in oracle/og_lsq.py -- what `scipy.optimize.minimize(method="SLSQP")` runs for the reference's
`Problem.solve` (/root/reference/OpenGoddard/optimize.py:738-755; driver scipy/optimize/_slsqp_py.py:453-560).
Restated from D. Kraft, "A software package for sequential quadratic programming" (DFVLR-FB 88-28, 1988) and the
state SciPy's driver exposes (`acc, alpha, f0, gs, h1..h4, t, t0, tol, reset, line, inconsistent`):

  * B = L D L' (BFGS), reset to the identity at most 5 times; QP by `og_lsq.lsq`,
  * the augmented QP (one slack variable in [0, 1], weight 100, x10 up to 5 times) when the linearised
    constraints are inconsistent (mode 4),
  * L1 penalty weights  mu_j = max(|r_j|, (mu_j + |r_j|) / 2),
  * convergence tests before (|g's| + sum |r_j c_j|, violation) and after the step (|f - f0|, |s|, violation),
    the relaxed test (10 acc) after 5 resets,
  * Armijo-type line search  alpha <- max(h3 / (2 (h3 - h1)), 0.1), at most 10 trials,
  * damped BFGS update (Powell) as two rank-one modifications of L D L' (Fletcher & Powell composite t).

Pinned against the installed SciPy (tests/test_device_sqp.py): on well-conditioned problems the iterates agree with
`minimize(method="SLSQP")` to rounding (same nit / nfev / status).  On the collocation problems SciPy >= 1.16
resolves the ill-conditioned first subproblems differently (its compiled LSQ core is not this algorithm's
Lawson-Hanson NNLS any more), so there only outcomes are compared.  This file is the checker of the device
solver (`csrc/ogb_sqp.h`); only tests/, smoke() and bench.py's CPU legs may import it."""
import numpy as np

from oracle import og_lsq

EPS = np.finfo(float).eps
ALFMIN = 0.1


def ldl_update(L, Dg, z, sigma):
    """L diag(Dg) L' + sigma z z', in place (SLSQP's `ldl`)."""
    n = len(z)
    if sigma == 0.0:
        return
    z = z.copy()
    t = 1.0 / sigma
    w = None
    if sigma < 0.0:
        w = z.copy()
        for i in range(n):
            v = w[i]
            t += v * v / Dg[i]
            w[i + 1:] -= v * L[i + 1:, i]
        if t >= 0.0:
            t = EPS / sigma
        tw = np.zeros(n)
        for i in range(n - 1, -1, -1):
            u = w[i]
            tw[i] = t
            t -= u * u / Dg[i]
        w = tw
    for i in range(n):
        v = z[i]
        delta = v / Dg[i]
        tp = w[i] if sigma < 0.0 else t + delta * v
        alpha = tp / t
        Dg[i] = alpha * Dg[i]
        if i == n - 1:
            break
        beta = delta / tp
        if alpha > 4.0:
            gamma = t / tp
            col = L[i + 1:, i].copy()
            L[i + 1:, i] = gamma * col + beta * z[i + 1:]
            z[i + 1:] -= v * col
        else:
            z[i + 1:] -= v * L[i + 1:, i]
            L[i + 1:, i] += beta * z[i + 1:]
        t = tp


def violation(c, meq):
    return float(np.sum(np.maximum(-c, np.where(np.arange(len(c)) < meq, c, 0.0))))


def search_direction(L, Dg, g, A, c, meq, lo, hi):
    """The QP of one SLSQP iteration, with the augmented problem on inconsistency.  Returns s (n), the
    multipliers r (m), h4, badlin, mode (1 = ok, else SLSQP's exit mode)."""
    n, m = len(g), len(c)
    s, r, mode = og_lsq.lsq(L, Dg, g, A, c, meq, lo, hi)
    badlin = False
    h4 = 1.0
    if mode == 6 and n == meq:
        mode = 4
    if mode == 4:
        badlin = True
        rho = 100.0
        acol = np.where(np.arange(m) < meq, -c, np.maximum(-c, 0.0))
        La = np.eye(n + 1)
        La[:n, :n] = L
        for incons in range(6):
            xa, r, mode = og_lsq.lsq(La, np.concatenate([Dg, [rho * rho]]), np.concatenate([g, [0.0]]),
                                     np.hstack([A, acol[:, None]]), c, meq, np.concatenate([lo, [0.0]]),
                                     np.concatenate([hi, [1.0]]))
            h4 = 1.0 - xa[n]
            s = xa[:n].copy()
            if mode != 4:
                break
            rho *= 10.0
    return s, r, h4, badlin, mode


def slsqp_numpy(evalf, evalg, x0, xl, xu, meq, acc=1e-6, maxiter=100, trace=None):
    """evalf(x) -> f, c (m,);  evalg(x) -> g (n,), A (m, n).  xl / xu: +-inf for none.  Returns the fields of
    SciPy's OptimizeResult this path reports (x, fun, status, nit, nfev, njev)."""
    xl = np.asarray(xl, dtype=float)
    xu = np.asarray(xu, dtype=float)
    x = np.clip(np.asarray(x0, dtype=float), xl, xu)
    n = len(x)
    f, c = evalf(x)
    g, A = evalg(x)
    m = len(c)
    nfev = njev = 1
    tol = 10.0 * acc
    it, reset = 0, 1                                  # (label 110: the initialisation counts as the first reset)
    mu = np.zeros(m)
    s = np.zeros(n)
    L, Dg = np.eye(n), np.ones(n)
    f0 = f
    badlin = False

    def finish(mode):
        return {"x": x, "fun": f, "status": mode, "nit": it, "nfev": nfev, "njev": njev}

    def relaxed():
        ok = (abs(f - f0) < tol or np.sqrt(s @ s) < tol) and violation(c, meq) < tol and not badlin and f == f
        return finish(0 if ok else 8)

    while True:
        it += 1
        if it > maxiter:
            it = maxiter                              # (SciPy reports nit = maxiter on exit 9)
            return finish(9)
        s, r, h4, badlin, qmode = search_direction(L, Dg, g, A, c, meq, xl - x, xu - x)
        if qmode != 1:
            return finish(qmode)
        if trace is not None:
            trace.append((it, x.copy(), s.copy(), r.copy(), h4))
        v = g - A.T @ r
        f0 = f
        x0_ = x.copy()
        gs = float(g @ s)
        h1 = abs(gs)
        h2 = 0.0
        for j in range(m):
            h3 = c[j] if j < meq else 0.0
            h2 += max(-c[j], h3)
            h3 = abs(r[j])
            mu[j] = max(h3, (mu[j] + h3) / 2.0)
            h1 += h3 * abs(c[j])
        if h1 < acc and h2 < acc and not badlin and f == f:
            return finish(0)
        h1 = 0.0
        for j in range(m):
            h3 = c[j] if j < meq else 0.0
            h1 += mu[j] * max(-c[j], h3)
        t0 = f + h1
        h3 = gs - h1 * h4
        if h3 >= 0.0:                                  # no descent: reset the BFGS matrix, new iteration
            reset += 1
            if reset > 5:
                return relaxed()
            L, Dg = np.eye(n), np.ones(n)
            continue
        line = 0
        alpha = 1.0
        while True:
            line += 1
            h3 = alpha * h3
            s = alpha * s
            x = x0_ + s
            f, c = evalf(np.clip(x, xl, xu))
            nfev += 1
            t = f
            for j in range(m):
                h1 = c[j] if j < meq else 0.0
                t += mu[j] * max(-c[j], h1)
            h1 = t - t0
            if h1 <= h3 / 10.0 or line > 10:
                break
            alpha = max(h3 / (2.0 * (h3 - h1)), ALFMIN)
        if (abs(f - f0) < acc or np.sqrt(s @ s) < acc) and violation(c, meq) < acc and not badlin and f == f:
            return finish(0)
        g, A = evalg(x)
        njev += 1
        u = g - A.T @ r - v
        w = L @ (Dg * (L.T @ s))
        h1 = float(s @ u)
        h2 = float(s @ w)
        h3 = 0.2 * h2
        if h1 < h3:
            h4 = (h2 - h3) / (h2 - h1)
            h1 = h3
            u = h4 * u + (1.0 - h4) * w
        if h1 == 0.0 or h2 == 0.0:
            reset += 1
            if reset > 5:
                return relaxed()
            L, Dg = np.eye(n), np.ones(n)
            continue
        ldl_update(L, Dg, u, 1.0 / h1)
        ldl_update(L, Dg, w, -1.0 / h2)
