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


TESTED_SCIPY = ((1, 16), (1, 19))      # [lo, hi): releases whose private `_slsqplib.slsqp` calling convention
#                                        (state dict, workspace sizing of _slsqp_py.py:453-520) this file mirrors


def _low_level():
    """SciPy's one-step SLSQP entry point.  It is private API: the state-dict keys, the workspace
    formula in _Instance and the NaN-means-unbounded convention are copied from SciPy's own driver,
    so an untested SciPy release is refused (a changed workspace layout would overflow the C core's
    buffers) unless $OGB200_ALLOW_UNTESTED_SCIPY=1; tests/test_sqp.py checks a batch of one against
    `minimize(method='SLSQP')` bit for bit on the installed SciPy."""
    import os
    import scipy
    ver = tuple(int(t) for t in scipy.__version__.split(".")[:2] if t.isdigit())
    if not (TESTED_SCIPY[0] <= ver < TESTED_SCIPY[1]) and os.environ.get("OGB200_ALLOW_UNTESTED_SCIPY") != "1":
        raise NotImplementedError("the batched SQP driver mirrors the private SLSQP step of SciPy %d.%d-%d.%d; "
                                  "installed: %s (set OGB200_ALLOW_UNTESTED_SCIPY=1 after running "
                                  "tests/test_sqp.py)" % (TESTED_SCIPY[0] + (TESTED_SCIPY[1][0], TESTED_SCIPY[1][1] - 1)
                                                          + (scipy.__version__,)))
    try:
        from scipy.optimize._slsqplib import slsqp
        from scipy.linalg.lapack import HAS_ILP64
    except Exception as e:                       # pragma: no cover
        raise NotImplementedError("the batched SQP driver needs SciPy's low-level SLSQP step "
                                  "(scipy >= 1.16): %r" % (e,))
    return slsqp, HAS_ILP64


class _Instance:
    """SLSQP state of one problem instance (mirrors scipy/optimize/_slsqp_py.py:453-520)."""

    def __init__(self, x0, n, m, meq, acc, maxiter, int_dtype, C=None, g=None):
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
        # the constraint normals SLSQP reads (scipy/optimize/_slsqp_py.py:517); its core never writes them,
        # so the zero background survives and an evaluator may rewrite only the structural non-zeros
        self.C = C if C is not None else np.zeros([max(1, m), n], dtype=np.float64, order="F")
        self.d = np.zeros([max(1, m)], dtype=np.float64)
        self.g = g if g is not None else np.zeros(n, dtype=np.float64)
        self.fx = 0.0
        self.nfev = self.njev = 0

    @property
    def mode(self):
        return self.state["mode"]


class _LocalStepper:
    """All SLSQP states in this process; the QP cores are stepped one after the other (SciPy's
    low-level step holds the GIL, so Python threads do not overlap them)."""

    def __init__(self, X0, n, m, meq, acc, maxiter, xl, xu, threads=1):
        slsqp, ilp64 = _low_level()
        self._slsqp, self.m, self.xl, self.xu = slsqp, m, xl, xu
        self.inst = [_Instance(x, n, m, meq, acc, maxiter, np.int64 if ilp64 else np.int32) for x in X0]
        self.pool = ThreadPoolExecutor(threads) if threads > 1 else None

    def put_values(self, ids, c):
        m = self.m
        for k, b in enumerate(ids):
            it = self.inst[b]
            it.fx = float(c[k, m])
            it.d[:m] = c[k, :m]

    def put_normals(self, ids, J, G):
        m = self.m
        for k, b in enumerate(ids):
            it = self.inst[b]
            it.C[:m, :] = J[k, :, :m].T
            it.g[:] = G[k] if G is not None else J[k, :, m]

    def _step(self, b):
        it = self.inst[b]
        self._slsqp(it.state, it.fx, it.g, it.C, it.d, it.x, it.mult, self.xl, self.xu, it.buffer, it.indices)

    def normals_buffer(self, ids):
        return None                  # (the per-instance C matrices are filled from the evaluator's array)

    def normals_targets(self, ids):
        """Addresses of the instances' own C matrices (Fortran order, leading dimension max(1, m)) and
        gradient vectors, for an evaluator that scatters the Jacobian's non-zeros straight into them."""
        return ([self.inst[b].C.ctypes.data for b in ids], [self.inst[b].g.ctypes.data for b in ids],
                max(1, self.m))

    def normals_written(self, ids, G):
        if G is not None:
            for k, b in enumerate(ids):
                self.inst[b].g[:] = G[k]

    def step(self, active):
        if self.pool is not None:
            list(self.pool.map(self._step, active))
        else:
            for b in active:
                self._step(b)

    def x(self, b):
        return self.inst[b].x

    def mode(self, b):
        return self.inst[b].state["mode"]

    def iters(self, b):
        return self.inst[b].state["iter"]

    def fx(self, b):
        return self.inst[b].fx

    def close(self):
        if self.pool is not None:
            self.pool.shutdown()


def _worker_loop(conn):
    """Body of a worker process (entry point: python -m opengoddard_b200.sqp_worker).  Serves any
    number of batched solves: every ("init", ...) message attaches the parent's shared memory and
    builds the SLSQP states of the instances this worker owns; ("step", ...) messages advance them."""
    from multiprocessing import shared_memory
    slsqp, ilp64 = _low_level()
    shms, arr, inst = {}, {}, {}

    def release():
        arr.clear()
        inst.clear()
        for sh in shms.values():
            sh.close()
        shms.clear()

    try:
        while True:
            msg = conn.recv()
            if msg[0] == "stop":
                break
            if msg[0] == "init":
                release()
                _, names, shapes, owned, n, m, meq, acc, maxiter, xl, xu = msg
                for k, v in names.items():
                    shms[k] = shared_memory.SharedMemory(name=v)
                try:                               # the parent owns (and unlinks) the segments: attaching must
                    from multiprocessing import resource_tracker     # not register them with this tracker (< 3.13)
                    for sh in shms.values():
                        resource_tracker.unregister(sh._name, "shared_memory")
                except Exception:
                    pass
                for k in names:
                    arr[k] = np.ndarray(shapes[k][0], dtype=shapes[k][1], buffer=shms[k].buf)
                for b in owned:       # C and g live in the parent's shared memory: nothing is copied per step
                    Cb = arr["C"][b].T if m > 0 else None              # (n, m) C-order block = (m, n) Fortran matrix
                    inst[b] = _Instance(arr["X"][b], n, m, meq, acc, maxiter, np.int64 if ilp64 else np.int32,
                                        C=Cb, g=arr["G"][b])
                conn.send("ready")
                continue
            if msg[0] == "release":
                release()
                conn.send("released")
                continue
            _, vals, active = msg
            for b in vals:
                it = inst[b]
                it.fx = float(arr["D"][b, m])
                it.d[:m] = arr["D"][b, :m]
            for b in active:
                it = inst[b]
                slsqp(it.state, it.fx, it.g, it.C, it.d, it.x, it.mult, xl, xu, it.buffer, it.indices)
                arr["X"][b] = it.x
                arr["mode"][b] = it.state["mode"]
                arr["iter"][b] = it.state["iter"]
                arr["fx"][b] = it.fx
            conn.send("done")
    finally:
        release()


class WorkerPool:
    """Worker processes for the SLSQP cores, reusable across slsqp_batch calls (Problem.solve_batch keeps
    one for all its outer passes).  The workers are plain `python -m opengoddard_b200.sqp_worker`
    subprocesses connected over a local socket -- not multiprocessing children -- so an unguarded
    user script (the reference's examples have no `if __name__ == "__main__"`) is never re-executed
    and no CUDA context is forked.  Each worker runs SLSQP's LAPACK with one BLAS thread."""

    def __init__(self, processes):
        import os
        import secrets
        import subprocess
        import sys
        import tempfile
        from multiprocessing import connection
        self.W = max(1, int(processes))
        self.conns, self.procs = [], []
        self._dir = tempfile.mkdtemp(prefix="ogb200_sqp_")
        try:
            key = secrets.token_bytes(16)
            address = os.path.join(self._dir, "sock")
            root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
            # one BLAS thread per worker: W workers already fill the cores.  (SLSQP's LAPACK calls round
            # differently with a threaded BLAS, so the result is bitwise the one a single process
            # produces under OMP_NUM_THREADS=1 / threadpoolctl.threadpool_limits(1).)
            env = dict(os.environ, OGB200_SQP_KEY=key.hex(), OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1",
                       MKL_NUM_THREADS="1", PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
            with connection.Listener(address, family="AF_UNIX", authkey=key) as listener:
                for w in range(self.W):
                    self.procs.append(subprocess.Popen([sys.executable, "-m", "opengoddard_b200.sqp_worker", address],
                                                       env=env, stdin=subprocess.DEVNULL))
                try:                                   # do not wait for ever for a worker that failed to start
                    listener._listener._socket.settimeout(120.0)
                except AttributeError:                 # (private attributes of multiprocessing.connection)
                    pass
                for w in range(self.W):
                    self.conns.append(listener.accept())
        except Exception:
            self.close()
            raise

    def wait(self, w, what, timeout=600.0):
        conn, waited = self.conns[w], 0.0
        while not conn.poll(0.5):
            waited += 0.5
            if any(pr.poll() is not None for pr in self.procs):
                raise RuntimeError("an SLSQP worker process exited unexpectedly")
            if waited >= timeout:
                raise RuntimeError("SLSQP worker process %d did not answer within %g s" % (w, timeout))
        if conn.recv() != what:
            raise RuntimeError("unexpected answer from SLSQP worker process %d" % w)

    def close(self):
        import shutil
        for conn in self.conns:
            try:
                conn.send(("stop",))
                conn.close()
            except Exception:
                pass
        for pr in self.procs:
            try:
                pr.wait(timeout=10)
            except Exception:
                pr.kill()
        self.conns, self.procs = [], []
        shutil.rmtree(self._dir, ignore_errors=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class _ProcessStepper:
    """The SLSQP states are spread over the processes of a WorkerPool (instance b belongs to worker
    b mod W), so the QP cores of a lock-step round run in parallel on the host cores; decision vectors,
    constraint values and Jacobians are exchanged through shared memory.  Same arithmetic per
    instance as _LocalStepper (the instances are independent)."""

    def __init__(self, X0, n, m, meq, acc, maxiter, xl, xu, pool, own_pool):
        from multiprocessing import shared_memory
        B = len(X0)
        self.m, self.B, self.pool, self.own_pool = m, B, pool, own_pool
        self.W = min(pool.W, B)
        # "C": per instance the (m, n) Fortran-ordered constraint normals SLSQP reads, i.e. an (n, m)
        # C-ordered block; "G": the cost gradients.  The workers' SLSQP states use these blocks in place.
        shapes = {"X": ((B, n), np.float64), "D": ((B, m + 1), np.float64), "C": ((B, n, max(1, m)), np.float64),
                  "G": ((B, n), np.float64), "mode": ((B,), np.int64), "iter": ((B,), np.int64),
                  "fx": ((B,), np.float64)}
        self.shms, self.arr = {}, {}
        self._vals = []
        try:
            for k, (shape, dt) in shapes.items():
                nbytes = max(8, int(np.prod(shape)) * np.dtype(dt).itemsize)
                self.shms[k] = shared_memory.SharedMemory(create=True, size=nbytes)
                self.arr[k] = np.ndarray(shape, dtype=dt, buffer=self.shms[k].buf)
                self.arr[k][...] = 0
            self.arr["X"][...] = X0
            names = {k: v.name for k, v in self.shms.items()}
            for w in range(self.W):
                pool.conns[w].send(("init", names, shapes, list(range(w, B, self.W)), n, m, meq, acc, maxiter, xl, xu))
            for w in range(self.W):
                pool.wait(w, "ready")
        except Exception:
            self.close()
            raise

    def put_values(self, ids, c):
        self.arr["D"][np.asarray(ids, dtype=np.int64)] = c
        self._vals = list(ids)

    def normals_buffer(self, ids):
        return None

    def normals_targets(self, ids):
        """Addresses of the instances' C blocks / gradient rows in shared memory (see _LocalStepper)."""
        C0, G0 = self.arr["C"].ctypes.data, self.arr["G"].ctypes.data
        cs, gs = self.arr["C"].strides[0], self.arr["G"].strides[0]
        return [C0 + b * cs for b in ids], [G0 + b * gs for b in ids], max(1, self.m)

    def normals_written(self, ids, G):
        if G is not None:
            self.arr["G"][np.asarray(ids, dtype=np.int64)] = G

    def put_normals(self, ids, J, G):
        m = self.m
        idx = np.asarray(ids, dtype=np.int64)
        if m > 0:
            self.arr["C"][idx] = J[:, :, :m]
        self.arr["G"][idx] = G if G is not None else J[:, :, m]

    def step(self, active):
        W = self.W
        for w in range(W):
            pick = lambda ids: [b for b in ids if b % W == w]
            self.pool.conns[w].send(("step", pick(self._vals), pick(active)))
        for w in range(W):
            self.pool.wait(w, "done")
        self._vals = []

    def x(self, b):
        return self.arr["X"][b]

    def mode(self, b):
        return int(self.arr["mode"][b])

    def iters(self, b):
        return int(self.arr["iter"][b])

    def fx(self, b):
        return float(self.arr["fx"][b])

    def close(self):
        try:
            if self.pool.conns and not self.own_pool:      # detach the workers before the segments go away
                for w in range(self.W):
                    self.pool.conns[w].send(("release",))
                for w in range(self.W):
                    self.pool.wait(w, "released", timeout=30.0)
        except Exception:
            pass
        if self.own_pool:
            self.pool.close()
        self.arr = {}
        for sh in self.shms.values():
            try:
                sh.close()
                sh.unlink()
            except Exception:
                pass
        self.shms = {}


def slsqp_batch(evaluator, X0, lb, ub, meq, mineq, ftol=1e-6, maxiter=25, cost_grad=None,
                threads=1, callback=None, processes=0):
    """Run SLSQP on every row of X0 in lock step.

    evaluator : object with eval(X) and eval_fd(X) (see module docstring)
    lb, ub    : (n,) bounds with +-inf for "none" (SciPy clips x0 into them first)
    cost_grad : optional callable x -> (n,) user gradient of the cost (reference
                `cost_derivative`, optimize.py:730-733); default = the FD row of J
    threads   : host threads stepping the per-instance QP cores (SciPy's step holds the GIL: no gain)
    processes : > 1 (or a WorkerPool to reuse): the per-instance SLSQP states live in that many worker
                processes and a lock-step
                round steps them in parallel (shared-memory exchange); bitwise the result of a single
                process running with one BLAS thread
    Returns dict(x (B, n), fun (B,), status (B,), nit (B,), nfev, njev, message list).
    """
    _low_level()
    X0 = np.atleast_2d(np.asarray(X0, dtype=np.float64))
    B, n = X0.shape
    m = int(meq + mineq)
    lb = np.asarray(lb, dtype=np.float64)
    ub = np.asarray(ub, dtype=np.float64)
    X0 = np.clip(X0, lb, ub)                                 # _slsqp_py.py:322
    xl = np.where(np.isfinite(lb), lb, np.nan)               # the C core wants NaN for "no bound"
    xu = np.where(np.isfinite(ub), ub, np.nan)
    if isinstance(processes, WorkerPool) and B >= 1:     # (also a batch of one: same BLAS threading as the rest)
        st = _ProcessStepper(X0, n, m, int(meq), float(ftol), maxiter, xl, xu, processes, False)
    elif not isinstance(processes, WorkerPool) and processes and processes > 1 and B > 1:
        st = _ProcessStepper(X0, n, m, int(meq), float(ftol), maxiter, xl, xu, WorkerPool(min(int(processes), B)), True)
    else:
        st = _LocalStepper(X0, n, m, int(meq), float(ftol), maxiter, xl, xu, threads)
    nfev = np.zeros(B, dtype=int)
    njev = np.zeros(B, dtype=int)

    def grads(ids):
        if cost_grad is None:
            return None
        return np.stack([np.asarray(cost_grad(np.array(st.x(b))), dtype=np.float64) for b in ids])

    try:
        # mode 0 on entry: objective, constraints and gradients at the start point
        ids = list(range(B))
        scatter = getattr(evaluator, "eval_fd_scatter", None)

        def normals(ids):
            """constraint values + Jacobians of the instances `ids` at their current x, into the SLSQP states:
            scattered by the evaluator straight into the instances' C / g buffers when it can, else copied
            from its dense (k, n, m + 1) array"""
            X = np.stack([st.x(b) for b in ids])
            G = grads(ids)
            if scatter is not None:
                Cp, Gp, ld = st.normals_targets(ids)
                c = scatter(X, Cp, ld, m, Gp if G is None else None)
                st.normals_written(ids, G)
            else:
                c, J = evaluator.eval_fd(X)
                st.put_normals(ids, J, G)
            return c

        st.put_values(ids, normals(ids))
        nfev += 1
        njev += 1
        active = list(range(B))
        iters_prev = [0] * B
        while active:
            st.step(active)
            need_f = [b for b in active if st.mode(b) == 1]
            need_g = [b for b in active if st.mode(b) == -1]
            if need_f:
                # SciPy clips x for the objective only (_clip_x_for_func); x stays inside the
                # bounds in exact arithmetic, so the clip matters for 1-2 ulp excursions
                Xf = np.stack([np.clip(st.x(b), lb, ub) for b in need_f])
                st.put_values(need_f, evaluator.eval(Xf))
                nfev[need_f] += 1
            if need_g:
                normals(need_g)
                njev[need_g] += 1
            if callback is not None:
                for b in active:
                    if st.iters(b) > iters_prev[b]:
                        callback(b, np.array(st.x(b)), st.fx(b))
                    iters_prev[b] = st.iters(b)
            active = [b for b in active if abs(st.mode(b)) == 1]
        modes = [st.mode(b) for b in range(B)]
        return {"x": np.stack([np.array(st.x(b)) for b in range(B)]),
                "fun": np.array([st.fx(b) for b in range(B)]),
                "status": np.array(modes), "nit": np.array([st.iters(b) for b in range(B)]),
                "nfev": nfev, "njev": njev, "message": [EXIT_MODES.get(k, "?") for k in modes]}
    finally:
        st.close()
