"""TEST INFRASTRUCTURE (oracle): numpy restatement of the least-squares core of Kraft's SLSQP -- LSQ -> LSEI ->
LSI -> LDP -> NNLS with Householder (H12) and Givens transformations, after D. Kraft, "A software package for
sequential quadratic programming" (DFVLR-FB 88-28, 1988) and C. L. Lawson & R. J. Hanson, "Solving Least Squares
Problems" (1974), chapters 10, 20, 23 -- the algorithm behind `scipy.optimize.minimize(method="SLSQP")`, which
is all the reference's `Problem.solve` calls (/root/reference/OpenGoddard/optimize.py:738-755; the reference pins
`scipy>=0.18`, setup.py:88: the Fortran `slsqp_optmz.f` of every SciPy up to 1.15, rewritten in C in 1.16).  SciPy's
compiled core is not in the reference tree; this file restates the published algorithm and is pinned against the
installed SciPy's own low-level step (tests/test_device_sqp.py).  Only tests/, smoke() and bench.py's CPU legs
may import it."""
import numpy as np

EPMACH = np.finfo(float).eps


def h12_construct(u, p, l1):
    """Householder H = I + u u'/b zeroing u[l1:], pivot p (Lawson & Hanson H12, mode 1).  `u` is modified
    (u[p] <- the new pivot value); returns `up` (None = identity)."""
    m = len(u)
    if not (0 <= p < l1 <= m - 1 + 0) or l1 > m - 1:
        return None
    cl = max(abs(u[p]), np.abs(u[l1:]).max())
    if cl <= 0.0:
        return None
    clinv = 1.0 / cl
    sm = (u[p] * clinv) ** 2 + float(np.sum((u[l1:] * clinv) ** 2))
    cl = cl * np.sqrt(sm)
    if u[p] > 0.0:
        cl = -cl
    up = u[p] - cl
    u[p] = cl
    return up


def h12_apply(u, p, l1, up, C):
    """Apply the transformation to the ROWS of C (each row c: c <- c + (c.v / b) v, v = (up at p, u[l1:]))."""
    if up is None or C.shape[0] == 0:
        return
    b = up * u[p]
    if b >= 0.0:
        return
    b = 1.0 / b
    sm = (C[:, p] * up + C[:, l1:] @ u[l1:]) * b
    C[:, p] += sm * up
    C[:, l1:] += np.outer(sm, u[l1:])


def nnls(A, b, itmax=None):
    """min ||A x - b||, x >= 0 (Lawson & Hanson NNLS).  A (m, n) and b are overwritten.  Returns x, rnorm, w
    (dual), mode (1 ok, 3 iteration limit)."""
    m, n = A.shape
    x = np.zeros(n)
    w = np.zeros(n)
    z = np.zeros(m)
    index = list(range(n))
    iz1, iz2 = 0, n - 1
    nsetp = 0
    npp1 = 0                                    # 0-based row of the next pivot (= nsetp)
    itmax = itmax or 3 * n
    it = 0
    mode = 1
    factor = 1.0e-2
    up_store = 0.0
    while True:
        # ---- loop A
        if iz1 > iz2 or nsetp >= m:
            break
        for iz in range(iz1, iz2 + 1):
            j = index[iz]
            w[j] = A[npp1:, j] @ b[npp1:]
        found = False
        while True:
            wmax, izmax = 0.0, -1
            for iz in range(iz1, iz2 + 1):
                j = index[iz]
                if w[j] > wmax:
                    wmax, izmax = w[j], iz
            if wmax <= 0.0:
                break
            iz = izmax
            j = index[iz]
            asave = A[npp1, j]
            col = A[:, j]
            up = h12_construct(col, npp1, npp1 + 1)
            unorm = np.sqrt(float(col[:nsetp] @ col[:nsetp]))
            t = factor * abs(A[npp1, j])
            if (unorm + t) - unorm > 0.0:
                z[:] = b
                h12_apply(col, npp1, npp1 + 1, up, z[None, :])
                if z[npp1] / A[npp1, j] > 0.0:
                    found = True
                    break
            A[npp1, j] = asave
            w[j] = 0.0
        if not found:
            break
        # ---- step 5: add column j
        b[:] = z
        index[iz] = index[iz1]
        index[iz1] = j
        iz1 += 1
        nsetp = npp1 + 1
        npp1 += 1
        if iz1 <= iz2:
            cols = [index[jz] for jz in range(iz1, iz2 + 1)]
            sub = np.ascontiguousarray(A[:, cols].T)
            h12_apply(A[:, j], nsetp - 1, npp1, up, sub)
            A[:, cols] = sub.T
        w[j] = 0.0
        A[npp1:, j] = 0.0
        # ---- loop B
        while True:
            for ip in range(nsetp - 1, -1, -1):
                if ip != nsetp - 1:
                    z[:ip + 1] -= z[ip + 1] * A[:ip + 1, jj]
                jj = index[ip]
                z[ip] = z[ip] / A[ip, jj]
            it += 1
            if it > itmax:
                mode = 3
                break
            alpha, jj = 1.0, -1
            for ip in range(nsetp):
                if z[ip] > 0.0:
                    continue
                l = index[ip]
                t = -x[l] / (z[ip] - x[l])
                if alpha < t:
                    continue
                alpha, jj = t, ip
            for ip in range(nsetp):
                l = index[ip]
                x[l] = (1.0 - alpha) * x[l] + alpha * z[ip]
            if jj < 0:
                break
            # ---- step 11: delete column
            i = index[jj]
            while True:
                x[i] = 0.0
                jj += 1
                for jcol in range(jj, nsetp):
                    ii = index[jcol]
                    index[jcol - 1] = ii
                    a, bb = A[jcol - 1, ii], A[jcol, ii]
                    roe = a if abs(a) > abs(bb) else bb          # (BLAS drotg)
                    scale = abs(a) + abs(bb)
                    if scale == 0.0:
                        c, s, sig = 1.0, 0.0, 0.0
                    else:
                        sig = scale * np.sqrt((a / scale) ** 2 + (bb / scale) ** 2)
                        sig = np.copysign(sig, roe)
                        c, s = a / sig, bb / sig
                    t = sig
                    r1, r2 = A[jcol - 1, :].copy(), A[jcol, :].copy()
                    A[jcol - 1, :] = c * r1 + s * r2
                    A[jcol, :] = -s * r1 + c * r2
                    A[jcol - 1, ii] = t
                    A[jcol, ii] = 0.0
                    b1, b2 = b[jcol - 1], b[jcol]
                    b[jcol - 1] = c * b1 + s * b2
                    b[jcol] = -s * b1 + c * b2
                npp1 = nsetp - 1
                nsetp -= 1
                iz1 -= 1
                index[iz1] = i
                if nsetp <= 0:
                    mode = 3
                    break
                again = False
                for jq in range(nsetp):
                    i = index[jq]
                    if x[i] <= 0.0:
                        jj = jq
                        again = True
                        break
                if not again:
                    break
            if mode != 1:
                break
            z[:] = b
        if mode != 1:
            break
    k = min(npp1, m - 1)
    rnorm = np.sqrt(float(b[k:] @ b[k:])) if nsetp < m else 0.0
    if npp1 > m - 1:
        w[:] = 0.0
    return x, rnorm, w, mode


def ldp(G, h):
    """min ||x|| s.t. G x >= h.  Returns x, xnorm, multipliers w (mg), mode (1 ok, 4 inconsistent, 3)."""
    mg, n = G.shape
    x = np.zeros(n)
    if mg == 0:
        return x, 0.0, np.zeros(0), 1
    E = np.empty((n + 1, mg))
    E[:n] = G.T
    E[n] = h
    f = np.zeros(n + 1)
    f[n] = 1.0
    u, rnorm, _, mode = nnls(E, f)
    if mode != 1:
        return x, 0.0, np.zeros(mg), mode
    if rnorm <= 0.0:
        return x, 0.0, np.zeros(mg), 4
    fac = 1.0 - float(h @ u)
    if not ((1.0 + fac) - 1.0 > 0.0):
        return x, 0.0, np.zeros(mg), 4
    fac = 1.0 / fac
    x = fac * (G.T @ u)
    return x, np.sqrt(float(x @ x)), fac * u, 1


def lsi(E, f, G, h):
    """min ||E x - f|| s.t. G x >= h (E (me, n) full column rank).  E, f, G, h are overwritten."""
    me, n = E.shape
    mg = G.shape[0]
    for i in range(n):
        col = E[:, i]
        up = h12_construct(col, i, i + 1) if i + 1 < me else None
        if up is not None:
            rest = np.ascontiguousarray(E[:, i + 1:].T)
            h12_apply(col, i, i + 1, up, rest)
            E[:, i + 1:] = rest.T
            h12_apply(col, i, i + 1, up, f[None, :])
    for j in range(n):
        if not (abs(E[j, j]) >= EPMACH):
            return np.zeros(n), 0.0, np.zeros(mg), 5
        G[:, j] = (G[:, j] - G[:, :j] @ E[:j, j]) / E[j, j]
    h -= G @ f[:n]
    x, xnorm, w, mode = ldp(G, h)
    if mode != 1:
        return x, xnorm, w, mode
    x = x + f[:n]
    for i in range(n - 1, -1, -1):
        x[i] = (x[i] - E[i, i + 1:n] @ x[i + 1:]) / E[i, i]
    t = np.sqrt(float(f[n:] @ f[n:])) if me > n else 0.0
    return x, np.sqrt(xnorm * xnorm + t * t), w, 1


def lsei(C, d, E, f, G, h):
    """min ||E x - f|| s.t. C x = d, G x >= h.  Returns x, multipliers w (mc + mg), mode.  Inputs overwritten."""
    mc, n = C.shape
    me, mg = E.shape[0], G.shape[0]
    if mc > n:
        return np.zeros(n), np.zeros(mc + mg), 2
    l = n - mc
    ups = []
    for i in range(mc):
        row = C[i, :]
        up = h12_construct(row, i, i + 1) if i + 1 < n else None
        ups.append(up)
        h12_apply(row, i, i + 1, up, C[i + 1:, :])
        h12_apply(row, i, i + 1, up, E)
        h12_apply(row, i, i + 1, up, G)
    x = np.zeros(n)
    for i in range(mc):
        if abs(C[i, i]) < EPMACH:
            return x, np.zeros(mc + mg), 6
        x[i] = (d[i] - C[i, :i] @ x[:i]) / C[i, i]
    w = np.zeros(mc + mg)
    if mc < n:
        f2 = f - E[:, :mc] @ x[:mc]
        E2 = E[:, mc:].copy()
        G2 = G[:, mc:].copy()
        if mg == 0:
            sol, *_ = np.linalg.lstsq(E2, f2, rcond=None)            # (HFTI; not used by the collocation problems)
            x[mc:] = sol
            mode = 1
        else:
            h2 = h - G[:, :mc] @ x[:mc]
            x2, xnorm, wg, mode = lsi(E2, f2, G2, h2)
            x[mc:] = x2
            w[mc:] = wg
            if mode != 1:
                return x, w, mode
    fres = E @ x - f
    dd = E[:, :mc].T @ fres - G[:, :mc].T @ w[mc:]
    for i in range(mc - 1, -1, -1):
        h12_apply(C[i, :], i, i + 1, ups[i], x[None, :])
    for i in range(mc - 1, -1, -1):
        w[i] = (dd[i] - C[i + 1:mc, i] @ w[i + 1:mc]) / C[i, i]
    return x, w, 1


def lsq(L, Dg, g, A, b, meq, xl, xu):
    """SLSQP's QP as a least-squares problem: min ||D^1/2 L' x + D^-1/2 L^-1 g|| s.t. A[:meq] x + b[:meq] = 0,
    A[meq:] x + b[meq:] >= 0, xl <= x <= xu (NaN / inf = none).  Returns x, y (multipliers of the m constraints),
    mode."""
    n = len(g)
    m = len(b)
    E = (L * np.sqrt(Dg)[None, :]).T.copy()                    # upper triangular D^1/2 L'
    f = np.zeros(n)
    for i in range(n):
        f[i] = (g[i] - E[:i, i] @ f[:i]) / E[i, i]
    f = -f
    C = A[:meq].copy()
    d = -b[:meq].copy()
    lo_ok = np.isfinite(xl)
    hi_ok = np.isfinite(xu)
    G = np.vstack([A[meq:], np.eye(n)[lo_ok], -np.eye(n)[hi_ok]])
    h = np.concatenate([-b[meq:], xl[lo_ok], -xu[hi_ok]])
    x, w, mode = lsei(C, d, E, f, G, h)
    y = np.zeros(m)
    if mode == 1:
        y[:] = w[:m]
        x = np.minimum(np.maximum(x, np.where(lo_ok, xl, -np.inf)), np.where(hi_ok, xu, np.inf))
    return x, y, mode
