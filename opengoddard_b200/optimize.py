"""Drop-in facade: ``from OpenGoddard.optimize import Problem, Guess, Condition, Dynamics``.

Same public names, argument meaning, attribute set and error behaviour as the
reference module (/root/reference/OpenGoddard/optimize.py -- cited per method), so the
shipped example scripts run unchanged, but the per-SQP-iteration hot path
(constraint vector + dense forward-difference Jacobian, reference :670-715 driven by
scipy/optimize/_slsqp_py.py:353-367) is not evaluated in Python: `solve` traces the user
callbacks once (trace.py), lowers them to device tapes (tape.py) and hands SciPy's
SLSQP `fun` / `jac` callables that resolve to the sm_100a kernels of libogb200.so
(engine.py).  New, batched entry points: `Problem.compile`, `Problem.evaluate_batch`.

Backends: "cuda" (default; raises if the library or a GPU is missing or a callback cannot be
traced -- there is no silent fallback), "auto" (explicit opt-in: like "cuda", but a problem whose
callbacks the tracer refuses is solved on the host with a RuntimeWarning naming the reason) and "host",
an explicit opt-in (``prob.backend = "host"`` or
``OGB200_BACKEND=host``) that evaluates the callbacks eagerly in numpy exactly like
the reference does; it exists for BASELINE.json's configs[0] (a single CPU instance,
"plumbing, no GPU") and is never used by the batched hot-path API.
"""
import os

import numpy as np
from scipy import interpolate
from scipy import optimize

from . import trace as _T

__all__ = ["Problem", "Guess", "Condition", "Dynamics"]


def _plt():
    import matplotlib.pyplot as plt
    return plt


def _lgl(capi, N):
    """tau, w, D of the N-point LGL rule (reference optimize.py:183-213) for the constructor
    attributes: libogb200's host entry point (the code the device kernel runs).  Only under the
    explicit host backend ($OGB200_BACKEND=host: BASELINE configs[0], "plumbing, no GPU") a
    missing library is replaced by the same Newton iteration in numpy; the cuda backend has no
    fallback and raises."""
    try:
        return capi.lgl_host(N)
    except (capi.OgbError, OSError):
        if os.environ.get("OGB200_BACKEND") != "host":
            raise
    n = N - 1
    x = -np.cos(np.pi * np.arange(N) / n)
    P = np.zeros((N, N))
    for _ in range(100):                          # Newton on (1 - x^2) P'_n(x) with the recurrence
        P[:, 0], P[:, 1] = 1.0, x
        for k in range(2, N):
            P[:, k] = ((2 * k - 1) * x * P[:, k - 1] - (k - 1) * P[:, k - 2]) / k
        step = (x * P[:, n] - P[:, n - 1]) / (N * P[:, n])
        x = x - step
        if np.abs(step).max() < 1e-16:
            break
    x[0], x[-1] = -1.0, 1.0
    if N % 2:
        x[n // 2] = 0.0
    Pn = P[:, n]
    w = 2.0 / (N * n * Pn ** 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        D = Pn[:, None] / Pn[None, :] / (x[:, None] - x[None, :])
    D[np.arange(N), np.arange(N)] = 0.0
    D[0, 0], D[-1, -1] = -N * n / 4.0, N * n / 4.0
    return x, w, D


class Problem:
    """OpenGoddard Problem (reference optimize.py:38-880).

    Args:
        time_init (list of float): [t_start, t_knot..., t_final]
        nodes (list of int): LGL nodes per phase
        number_of_states (list of int), number_of_controls (list of int)
        maxIterator (int): outer restarts of SLSQP in `solve`
        method: ignored, as in the reference (only LGL is wired, :787-789)
    """

    def __init__(self, time_init, nodes, number_of_states, number_of_controls,
                 maxIterator=100, method="LGL"):
        assert isinstance(time_init, list), \
            "error: time_init is not list"
        assert isinstance(nodes, list), \
            "error: nodes are not list"
        assert isinstance(number_of_states, list), \
            "error: number of states are not list"
        assert isinstance(number_of_controls, list), \
            "error: number of controls are not list"
        assert len(time_init) == len(nodes) + 1, \
            "error: time_init length is not match nodes length"
        assert len(nodes) == len(number_of_states), \
            "error: nodes length is not match states length"
        assert len(nodes) == len(number_of_controls), \
            "error: nodes length is not match controls length"
        from . import capi
        self.nodes = nodes
        self.number_of_states = number_of_states
        self.number_of_controls = number_of_controls
        self.number_of_section = len(nodes)
        # end offset of every block, phase by phase (reference :237-245)
        self.div = []
        end = 0
        for N, ns, nc in zip(nodes, number_of_states, number_of_controls):
            self.div.append([end + N * (k + 1) for k in range(ns + nc)])
            end = self.div[-1][-1]
        self.number_of_param = np.array(number_of_states) + np.array(number_of_controls)
        self.number_of_variables = int(sum(self.number_of_param * nodes)) + self.number_of_section
        self.tau, self.w, self.D, self.time = [], [], [], []
        for i, N in enumerate(nodes):
            tau, w, D = _lgl(capi, N)
            self.tau.append(tau)
            self.w.append(w)
            self.D.append(D)
            self.time.append((time_init[i + 1] - time_init[i]) / 2.0 * tau
                             + (time_init[i + 1] + time_init[i]) / 2.0)
        self.maxIterator = maxIterator
        self.iterator = 0
        self.time_init = time_init
        self.t0 = time_init[0]
        self.time_all_section = np.concatenate([t for t in self.time])
        self.unit_states = [[1.0] * ns for ns in number_of_states]
        self.unit_controls = [[1.0] * nc for nc in number_of_controls]
        self.unit_time = 1.0
        self.p = np.zeros(self.number_of_variables, dtype=float)
        self.bounds = [(None, None)] * self.number_of_variables
        for s in range(self.number_of_section):
            self.set_time_final_bounds(s, 0.0, None)
        self.dynamics = []
        self.knot_states_smooth = []
        self.cost = None
        self.running_cost = None
        self.cost_derivative = None
        self.equality = None
        self.inequality = None
        for s in range(self.number_of_section):
            self.set_time_final(s, time_init[s + 1])
            self.dynamics.append(None)
        for s in range(self.number_of_section - 1):
            self.knot_states_smooth.append(True)
        # ---- B200 engine state (not in the reference)
        self.backend = None          # None -> $OGB200_BACKEND -> "cuda"
        self.device = None           # torch device for the cuda backend (default cuda:0)
        self._engine = None
        self._engine_key = None

    # ------------------------------------------------------------------ layout
    def _span_state(self, state, section):
        # reference :247-260 (python negative indices are deliberately not normalised:
        # states(-1, s) addresses the LAST block of the phase)
        assert section < len(self.nodes), \
            "section argument out of own section range"
        assert state < self.number_of_states[section], \
            "states argument out of own states range"
        if state == 0:
            lo = 0 if section == 0 else self.div[section - 1][-1]
        else:
            lo = self.div[section][state - 1]
        return lo, self.div[section][state]

    def _span_control(self, control, section):
        # reference :262-269
        assert section < len(self.nodes), \
            "section argument out of own section range"
        assert control < self.number_of_controls[section], \
            "controls argument out of own controls range"
        k = self.number_of_states[section] + control
        return self.div[section][k - 1], self.div[section][k]

    def _division_states(self, state, section):
        lo, hi = self._span_state(state, section)
        return hi, lo

    def _division_controls(self, control, section):
        lo, hi = self._span_control(control, section)
        return hi, lo

    # ------------------------------------------------------------------ getters (:271-375)
    def states(self, state, section):
        """dimensional values of one state over the nodes of one phase"""
        lo, hi = self._span_state(state, section)
        return self.p[lo:hi] * self.unit_states[section][state]

    def states_all_section(self, state):
        out = np.zeros(0)
        for s in range(self.number_of_section):
            out = np.concatenate([out, self.states(state, s)])
        return out

    def controls(self, control, section):
        lo, hi = self._span_control(control, section)
        return self.p[lo:hi] * self.unit_controls[section][control]

    def controls_all_section(self, control):
        out = np.zeros(0)
        for s in range(self.number_of_section):
            out = np.concatenate([out, self.controls(control, s)])
        return out

    def time_start(self, section):
        if section == 0:
            return self.t0
        return self.p[range(-self.number_of_section - 1, 0)[section]] * self.unit_time

    def time_final(self, section):
        return self.p[range(-self.number_of_section, 0)[section]] * self.unit_time

    def time_final_all_section(self):
        return [self.time_final(s) for s in range(self.number_of_section)]

    # ------------------------------------------------------------------ setters (:377-440)
    def set_states(self, state, section, value):
        assert len(value) == self.nodes[section], "Error: value length is NOT match nodes length"
        lo, hi = self._span_state(state, section)
        self.p[lo:hi] = value / self.unit_states[section][state]

    def set_states_all_section(self, state, value_all_section):
        at = 0
        for s in range(self.number_of_section):
            self.set_states(state, s, value_all_section[at:at + self.nodes[s]])
            at += self.nodes[s]

    def set_controls(self, control, section, value):
        assert len(value) == self.nodes[section], "Error: value length is NOT match nodes length"
        lo, hi = self._span_control(control, section)
        self.p[lo:hi] = value / self.unit_controls[section][control]

    def set_controls_all_section(self, control, value_all_section):
        at = 0
        for s in range(self.number_of_section):
            self.set_controls(control, s, value_all_section[at:at + self.nodes[s]])
            at += self.nodes[s]

    def set_time_final(self, section, value):
        self.p[range(-self.number_of_section, 0)[section]] = value / self.unit_time

    # ------------------------------------------------------------------ bounds (:442-507)
    def set_states_bounds(self, state, section, lb, ub):
        u = self.unit_states[section][state]
        lb = lb / u if lb is not None else None
        ub = ub / u if ub is not None else None
        lo, hi = self._span_state(state, section)
        self.bounds[lo:hi] = [(lb, ub)] * self.nodes[section]

    def set_states_bounds_all_section(self, state, lb, ub):
        for s in range(self.number_of_section):
            self.set_states_bounds(state, s, lb, ub)

    def set_controls_bounds(self, control, section, lb, ub):
        u = self.unit_controls[section][control]
        lb = lb / u if lb is not None else None
        ub = ub / u if ub is not None else None
        lo, hi = self._span_control(control, section)
        self.bounds[lo:hi] = [(lb, ub)] * self.nodes[section]

    def set_controls_bounds_all_section(self, control, lb, ub):
        for s in range(self.number_of_section):
            self.set_controls_bounds(control, s, lb, ub)

    def set_time_final_bounds(self, section, lb, ub):
        lb = lb / self.unit_time if lb is not None else 0.0
        ub = ub / self.unit_time if ub is not None else None
        self.bounds[self.index_time_final(section)] = (lb, ub)

    def bounds_arrays(self):
        """(lb, ub) float64 arrays with +-inf for None (scipy old_bound_to_new)."""
        lb = np.array([-np.inf if b[0] is None else float(b[0]) for b in self.bounds])
        ub = np.array([np.inf if b[1] is None else float(b[1]) for b in self.bounds])
        return lb, ub

    # ------------------------------------------------------------------ time helpers (:509-540)
    def time_to_tau(self, time):
        lo, hi = min(time), max(time)
        mid = (lo + hi) / 2
        return np.array([2 / (hi - lo) * (x - mid) for x in time])

    def time_update(self):
        self.time = []
        t = [0] + self.time_final_all_section()
        for i in range(self.number_of_section):
            self.time.append((t[i + 1] - t[i]) / 2.0 * self.tau[i] + (t[i + 1] + t[i]) / 2.0)
        return np.concatenate([i for i in self.time])

    def time_knots(self):
        return [0] + self.time_final_all_section()

    # ------------------------------------------------------------------ index helpers (:542-572)
    def index_states(self, state, section, index=None):
        lo, hi = self._span_state(state, section)
        if index is None:
            return lo
        assert index < hi - lo, "Error, index out of range"
        if index < 0:
            index = hi - lo + index
        return lo + index

    def index_controls(self, control, section, index=None):
        lo, hi = self._span_control(control, section)
        if index is None:
            return lo
        assert index < hi - lo, "Error, index out of range"
        if index < 0:
            index = hi - lo + index
        return lo + index

    def index_time_final(self, section):
        return self.number_of_variables + range(-self.number_of_section, 0)[section]

    # ------------------------------------------------------------------ unit scaling (:579-639)
    def set_unit_states(self, state, section, value):
        self.unit_states[section][state] = value

    def set_unit_states_all_section(self, state, value):
        for s in range(self.number_of_section):
            self.set_unit_states(state, s, value)

    def set_unit_controls(self, control, section, value):
        self.unit_controls[section][control] = value

    def set_unit_controls_all_section(self, control, value):
        for s in range(self.number_of_section):
            self.set_unit_controls(control, s, value)

    def set_unit_time(self, value):
        self.unit_time = value
        ti = np.array(self.time_init) / value
        self.time_init = list(ti)
        self.time = []
        for i in range(self.number_of_section):
            self.time.append((ti[i + 1] - ti[i]) / 2.0 * self.tau[i] + (ti[i + 1] + ti[i]) / 2.0)
        self.t0 = ti[0]
        self.time_all_section = np.concatenate([t for t in self.time])
        for s in range(self.number_of_section):
            self.set_time_final(s, ti[s + 1] * value)

    # ------------------------------------------------------------------ B200 engine
    def _backend_name(self):
        name = self.backend or os.environ.get("OGB200_BACKEND", "cuda")
        if name not in ("cuda", "host", "auto"):
            raise ValueError("backend must be 'cuda', 'host' or 'auto', not %r" % (name,))
        return name

    def _check_callbacks(self):
        assert len(self.dynamics) != 0, "It must be set dynamics"
        assert self.cost is not None, "It must be set cost function"
        assert self.equality is not None, "It must be set equality function"
        assert self.inequality is not None, "It must be set inequality function"

    def compile(self, obj, device=None, jit=True):
        """Trace the callbacks once and build the device engine (cuda backend only).
        Returns an `engine.DeviceProblem`; raises if libogb200.so or a GPU is missing.
        jit=True also compiles the traced tapes into the sweep kernel (NVRTC)."""
        from . import engine, tape
        self._check_callbacks()
        ir = tape.build_ir(self, obj)
        self._engine = engine.DeviceProblem(ir, self.bounds_arrays(), device or self.device, jit=jit)
        self._engine_key = self._fingerprint(obj, jit)
        return self._engine

    def _fingerprint(self, obj, jit):
        """Everything a compiled engine has baked in: the callbacks, the constants they read from
        `obj`, bounds (clipping and FD-step flipping happen on the device), units, knot flags, the
        layout and the kernel flavour.  `_engine_for` recompiles when it changes."""
        def freeze(v, depth=0):
            if isinstance(v, (bool, int, float, complex, str, bytes, type(None))):
                return ("v", repr(v))
            if isinstance(v, np.generic):
                return ("v", repr(v.item()))
            if isinstance(v, np.ndarray):
                return ("a", v.dtype.str, v.shape, v.tobytes())
            if isinstance(v, (list, tuple)) and depth < 4:
                return ("l",) + tuple(freeze(e, depth + 1) for e in v)
            if isinstance(v, dict) and depth < 4:
                return ("d",) + tuple((repr(k), freeze(e, depth + 1)) for k, e in sorted(v.items(), key=lambda kv: repr(kv[0])))
            return ("id", id(v))
        try:
            attrs = tuple((k, freeze(v)) for k, v in sorted(vars(obj).items()))
        except TypeError:
            attrs = ()
        lb, ub = self.bounds_arrays()
        cbs = tuple(id(f) for f in list(self.dynamics) + [self.cost, self.running_cost, self.equality,
                                                          self.inequality])
        return (id(obj), attrs, lb.tobytes(), ub.tobytes(), freeze(self.unit_states), freeze(self.unit_controls),
                float(self.unit_time), float(self.t0), freeze(list(self.knot_states_smooth)), tuple(self.nodes),
                tuple(self.number_of_states), tuple(self.number_of_controls), cbs, bool(jit))

    def _engine_for(self, obj, jit=True, device=None):
        """The cached engine if it was compiled for exactly this problem state, else a fresh one."""
        device = device or self.device
        if self._engine is not None and (device is None or str(self._engine.device) == str(device)) and \
                (obj is None or self._engine_key == self._fingerprint(obj, jit)):
            return self._engine
        if obj is None:
            raise ValueError("no compiled engine yet: pass `obj`")
        return self.compile(obj, device=device, jit=jit)

    def evaluate_batch(self, P, obj=None, jacobian=True, host=False, autotune=False):
        """Batched hot path: P (B, nvars) -> c (B, m+1) [, J (B, nvars, m+1)] as torch CUDA
        tensors; row m carries cost / grad cost.  J[b, j, :] is column j.  host=True: P is a host
        array and the results come back as numpy arrays in host memory through the host-buffer
        session (ogb_host_eval_fd: packed device->host transport, dense J rebuilt by host threads).
        jacobian: True = SciPy's forward differences (the reference's Jacobian), "sparse" = the same as
        packed values (B, nnz) in the engine's jac_pattern() layout, "exact" = the exact Jacobian (analytic
        collocation block + forward-mode tangents of the callbacks; opt-in, packed), False = c only.
        autotune=True (dense FD Jacobian on the device only): the first call on a compiled problem times the sweep
        kernel's CTA widths on P (DeviceProblem.autotune: the best width is problem-dependent, up to 18 % on the
        BASELINE problems) and keeps the choice for later calls; results do not depend on it."""
        eng = self._engine_for(obj)
        if autotune and jacobian is True and not host and getattr(eng, "tuned_threads", None) is None and len(P) > 0:
            eng.autotune(P)
        if jacobian in ("exact", "sparse"):
            if host:
                ev = eng.host_evaluator(exact=jacobian == "exact")
                ev._session_for(len(P))
                return ev.session.eval_fd(np.asarray(P, dtype=np.float64), mode="packed")
            return eng.eval_exact(P) if jacobian == "exact" else eng.eval_sparse(P)
        if host:
            ev = eng.host_evaluator()
            return ev.eval_fd(np.asarray(P, dtype=np.float64)) if jacobian else ev.eval(np.asarray(P, dtype=np.float64))
        return eng.eval_fd(P) if jacobian else eng.eval(P)

    def make_starts(self, B, obj=None, seed=20261017, first=0, rel_states=0.01, rel_time=0.05):
        """B multi-start decision vectors around the current guess `self.p`, generated ON THE DEVICE
        (engine.make_starts: counter-based Philox jitter, 1 % Gaussian on states / controls, +-5 % uniform on
        the final times, clipped into the bounds).  Returns a (B, nvars) CUDA tensor; instance `first + i` is
        the same whatever B or the number of GPUs."""
        eng = self._engine_for(obj)
        return eng.make_starts(self.p, B, seed, first, rel_states, rel_time)

    def guess_batch(self, specs, params, obj=None, tfinal=None):
        """A batch of guesses from per-instance boundary values, on the device (engine.guess_batch):
        specs = [("linear", ("state", 0), None), ("cubic", ("control", 0), 1), ...] name the generator, the
        variable and the phase (None = all phases, like Guess.*(prob.time_all_section) + set_*_all_section);
        params (B, len(specs), 4) hold const | (y0, yf) | (y0, y'0, yf, y'f) per instance."""
        eng = self._engine_for(obj)
        low = []
        for kind, (what, k), sec in specs:
            ns = self.number_of_states[0 if sec is None else sec]
            low.append((kind, k if what == "state" else ns + k, sec))
        return eng.guess_batch(low, params, self.time_all_section, tfinal=tfinal, base=self.p)

    def trajectories(self, P, obj=None):
        """time / states / controls of every instance of P at every node, dimensional, on the device:
        (B, total nodes, 1 + nstates + ncontrols) -- the batched time_update + *_all_section + to_csv table."""
        return self._engine_for(obj).trajectories(P)

    def solve_batch(self, P0, obj, ftol=1e-6, maxiter=25, max_outer=None, threads=1, group=None, processes=0,
                    jacobian="fd", qp="scipy"):
        """Multi-start: solve the NLP from every row of P0 (B, nvars) at once.

        Under an initialised torch.distributed process group (one rank per GPU) the rows of P0
        are sharded contiguously over the ranks, each rank solves its slice on its own GPU with no
        communication, and the per-instance results are all-gathered at the end (NCCL), so every
        rank returns the full result (batch.run_sharded).

        Every instance runs SciPy's SLSQP state machine (same C core as `solve`); all function
        and Jacobian evaluations of an SQP step are served by one batched device call
        (sqp.slsqp_batch).  Like `solve` (reference optimize.py:738-755) instances that did not
        reach exit mode 0 are restarted from where they stopped, up to `max_outer`
        (default maxIterator) times.  `processes` > 1 steps the per-instance SLSQP cores in that many
        worker processes (SciPy's step holds the GIL, so this is what makes the host side scale with
        the cores; sqp._ProcessStepper); a sqp.WorkerPool instance is used as is and left running, so
        repeated calls do not pay for starting the workers.  jacobian="exact": SLSQP is given the exact Jacobians (opt-in; the
        reference's are forward differences).  qp="device" (opt-in): the whole SLSQP iteration -- BFGS update, the
        QP, line search, convergence tests -- runs on the GPU too, one thread block per instance
        (engine.DeviceSqp, csrc/ogb_sqp.h: Kraft's SLSQP restated, not SciPy's compiled core; identical iterates on
        well-conditioned problems, same algorithm but not bitwise SciPy on ill-conditioned ones).
        Returns dict(x, fun, status, nit, outer)."""
        from . import batch, sqp
        self._check_callbacks()
        if qp not in ("scipy", "device"):
            raise ValueError("qp must be 'scipy' or 'device'")
        if qp == "device" and self.cost_derivative is not None:
            import warnings
            warnings.warn("OpenGoddard-B200: solve_batch(qp='device') takes the cost gradient from the device Jacobian "
                          "(finite differences, or exact with jacobian='exact'); the Python cost_derivative is not called",
                          RuntimeWarning, stacklevel=2)
        eng = self._engine_for(obj)
        P0 = np.array(np.atleast_2d(P0), dtype=np.float64)
        return batch.run_sharded(
            lambda rows: self._solve_rows(eng, rows, obj, ftol, maxiter, max_outer, threads, sqp, processes,
                                          exact=(jacobian == "exact"), device_qp=(qp == "device")),
            P0, group=group, device=eng.device)

    def _solve_rows(self, eng, P0, obj, ftol, maxiter, max_outer, threads, sqp, processes=0, exact=False,
                    device_qp=False):
        lb, ub = self.bounds_arrays()
        X = np.array(np.atleast_2d(P0), dtype=np.float64).reshape(-1, self.number_of_variables)
        B = X.shape[0]
        status = np.full(B, 9)
        fun = np.zeros(B)
        nit = np.zeros(B, dtype=int)
        outer = np.zeros(B, dtype=int)
        if B == 0:                                          # an empty shard (fewer starts than ranks)
            return {"x": X, "fun": fun, "status": status, "nit": nit, "outer": outer}
        if device_qp:
            with eng.device_sqp(B, ftol, maxiter) as dq:
                for _ in range(self.maxIterator if max_outer is None else max_outer):
                    ids = np.nonzero(status != 0)[0]
                    if ids.size == 0:
                        break
                    res = dq.solve(X[ids], exact=exact)
                    X[ids] = res["x"]
                    status[ids] = res["status"]
                    fun[ids] = res["fun"]
                    nit[ids] += res["nit"]
                    outer[ids] += 1
            return {"x": X, "fun": fun, "status": status, "nit": nit, "outer": outer}
        grad = None
        if self.cost_derivative is not None:
            def grad(x):
                self.p = x
                return np.ascontiguousarray(self.cost_derivative(self, obj), dtype=np.float64)
        own_pool = not isinstance(processes, sqp.WorkerPool)
        if own_pool:
            pool = sqp.WorkerPool(min(int(processes), B)) if processes and processes > 1 and B > 1 else None
        else:
            pool = processes                                # the caller's pool (kept alive across solve_batch calls)
        try:                                                # one set of worker processes for all outer passes
            for _ in range(self.maxIterator if max_outer is None else max_outer):
                ids = np.nonzero(status != 0)[0]
                if ids.size == 0:
                    break
                res = sqp.slsqp_batch(eng.host_evaluator(exact=exact), X[ids], lb, ub, eng.meq, eng.mineq, ftol=ftol,
                                      maxiter=maxiter, cost_grad=grad, threads=threads,
                                      processes=pool if pool is not None else 0)
                X[ids] = res["x"]
                status[ids] = res["status"]
                fun[ids] = res["fun"]
                nit[ids] += res["nit"]
                outer[ids] += 1
        finally:
            if pool is not None and own_pool:
                pool.close()
        return {"x": X, "fun": fun, "status": status, "nit": nit, "outer": outer}

    # ------------------------------------------------------------------ solve (:649-755)
    def _dummy_func():
        pass

    def solve(self, obj, display_func=_dummy_func, **options):
        """solve NLP with SciPy SLSQP; ftol (default 1e-6), maxiter (default 25)"""
        self._check_callbacks()
        backend = self._backend_name()
        if backend == "auto":
            # explicit, per-problem opt-in: callbacks the tracer cannot compile for the device (cross-node reads
            # inside dynamics, data-dependent python, ...) are evaluated eagerly in numpy like the reference does,
            # with a warning that says so.  A missing library or GPU still raises: this is not a silent CPU path.
            from . import trace
            try:
                fun, cons, jac = self._device_callables(obj)
            except trace.TraceError as ex:
                import warnings
                warnings.warn("OpenGoddard-B200: this problem's callbacks cannot be compiled for the device (%s); "
                              "backend='auto' evaluates them on the host (numpy, reference semantics) instead" % ex,
                              RuntimeWarning, stacklevel=2)
                self.fallback_reason = str(ex)
                fun, cons, jac = self._host_callables(obj)
        elif backend == "cuda":
            fun, cons, jac = self._device_callables(obj)
        else:
            fun, cons, jac = self._host_callables(obj)
        ftol = options.setdefault("ftol", 1e-6)
        maxiter = options.setdefault("maxiter", 25)
        # $OGB200_MAX_OUTER caps the outer restarts (test runs of the shipped scripts, which ask for up to 90)
        cap = int(os.environ.get("OGB200_MAX_OUTER", "0") or 0)
        first = self.iterator
        while self.iterator < self.maxIterator and (cap <= 0 or self.iterator - first < cap):
            print("---- iteration : {0} ----".format(self.iterator + 1))
            opt = optimize.minimize(fun, self.p, args=(self, obj), bounds=self.bounds,
                                    constraints=cons, jac=jac, method="SLSQP",
                                    options={"disp": True, "maxiter": maxiter, "ftol": ftol})
            print(opt.message)
            display_func()
            print("")
            if not (opt.status):
                break
            self.iterator += 1

    def _device_callables(self, obj):
        """SciPy-facing closures whose values AND Jacobians come from the CUDA kernels."""
        from . import engine
        eng = self._engine_for(obj, jit=False)   # one instance per call: latency-bound, skip NVRTC
        grad = None
        if self.cost_derivative is not None:
            def grad(x):                                  # user gradient, host (reference :733)
                self.p = x
                return self.cost_derivative(self, obj)

        def on_x(x):
            self.p = x                                    # reference for_solver, :713
        return engine.scipy_callables(eng, on_x=on_x, cost_derivative=grad, args=(self, obj))

    def _host_callables(self, obj):
        """Explicit opt-in numpy path (BASELINE.json configs[0]): the reference's closures."""
        def equality_add(p, *a):
            self.p = p
            rows = [np.atleast_1d(self.equality(self, obj))]
            for s in range(self.number_of_section):
                deriv = [self.D[s].dot(self.states(a_, s) / self.unit_states[s][a_])
                         for a_ in range(self.number_of_states[s])]
                tix = self.time_start(s) / self.unit_time
                tfx = self.time_final(s) / self.unit_time
                rows.append(np.concatenate(deriv) - (tfx - tix) / 2.0 * self.dynamics[s](self, obj, s))
            for k in range(self.number_of_section - 1):
                if self.number_of_states[k] != self.number_of_states[k + 1]:
                    continue
                for a_ in range(self.number_of_states[k]):
                    prev = self.states(a_, k) / self.unit_states[k][a_]
                    post = self.states(a_, k + 1) / self.unit_states[k][a_]
                    if self.knot_states_smooth[k]:
                        rows.append(np.atleast_1d(prev[-1] - post[0]))
            return np.concatenate(rows)

        def inequality(p, *a):
            self.p = p
            return self.inequality(self, obj)

        def cost_add(p, *a):
            self.p = p
            base = self.cost(self, obj)
            if self.running_cost is None:
                return base
            return base + sum(self.running_cost(self, obj) * np.concatenate([w for w in self.w]))

        cons = ({"type": "eq", "fun": equality_add, "args": (self, obj)},
                {"type": "ineq", "fun": inequality, "args": (self, obj)})
        jac = None
        if self.cost_derivative is not None:
            def jac(p, *a):
                self.p = p
                return self.cost_derivative(self, obj)
        return cost_add, cons, jac

    # ------------------------------------------------------------------ reporting (:825-880)
    def __repr__(self):
        s = "---- parameter ----" + "\n"
        s += "nodes = " + str(self.nodes) + "\n"
        s += "number of states    = " + str(self.number_of_states) + "\n"
        s += "number of controls  = " + str(self.number_of_controls) + "\n"
        s += "number of sections  = " + str(self.number_of_section) + "\n"
        s += "number of variables = " + str(self.number_of_variables) + "\n"
        s += "---- algorithm ----" + "\n"
        s += "max iteration = " + str(self.maxIterator) + "\n"
        s += "---- function  ----" + "\n"
        s += "dynamics        = " + str(self.dynamics) + "\n"
        s += "cost            = " + str(self.cost) + "\n"
        s += "cost_derivative = " + str(self.cost_derivative) + "\n"
        s += "equality        = " + str(self.equality) + "\n"
        s += "inequality      = " + str(self.inequality) + "\n"
        s += "knot_states_smooth = " + str(self.dynamics) + "\n"
        return s

    def to_csv(self, filename="OpenGoddard_output.csv", delimiter=","):
        cols = [self.time_update()]
        header = "time, "
        for i in range(self.number_of_states[0]):
            header += "state%d, " % (i)
            cols.append(self.states_all_section(i))
        for i in range(self.number_of_controls[0]):
            header += "control%d, " % (i)
            cols.append(self.controls_all_section(i))
        np.savetxt(filename, np.vstack(cols).T, delimiter=delimiter, header=header)
        print("Completed saving \"%s\"" % (filename))

    def plot(self, title_comment=""):
        plt = _plt()
        plt.figure()
        plt.title("OpenGoddard inner variables" + title_comment)
        plt.plot(self.p, "o")
        plt.xlabel("variables")
        plt.ylabel("value")
        for section in range(self.number_of_section):
            for line in self.div[section]:
                plt.axvline(line, color="C%d" % ((section + 1) % 6), alpha=0.5)
        plt.grid()


class Guess:
    """Initial-guess helpers (reference optimize.py:883-975)."""

    @classmethod
    def zeros(cls, time):
        return np.zeros(len(time))

    @classmethod
    def constant(cls, time, const):
        return np.ones(len(time)) * const

    @classmethod
    def linear(cls, time, y0, yf):
        f = interpolate.interp1d(np.array([time[0], time[-1]]), np.array([y0, yf]))
        return f(time)

    @classmethod
    def cubic(cls, time, y0, yprime0, yf, yprimef):
        t0, tf = time[0], time[-1]
        A = np.array([[1, t0, t0 ** 2, t0 ** 3], [0, 1, 2 * t0, 3 * t0 ** 2],
                      [1, tf, tf ** 2, tf ** 3], [0, 1, 2 * tf, 3 * tf ** 2]])
        C = np.linalg.inv(A).dot(np.array([y0, yprime0, yf, yprimef]))
        return C[0] + C[1] * time + C[2] * time ** 2 + C[3] * time ** 3

    @classmethod
    def plot(cls, x, y, title="", xlabel="", ylabel=""):
        plt = _plt()
        plt.figure()
        plt.plot(x, y, "-o")
        plt.title(title)
        plt.xlabel(xlabel)
        plt.ylabel(ylabel)
        plt.grid()


class Condition(object):
    """Row builder for user equality / inequality functions and dense cost gradients
    (reference optimize.py:978-1072).  While the callbacks are being traced the rows are
    symbolic and `__call__` returns them for the tape compiler."""

    def __init__(self, length=0):
        self._condition = np.zeros(length)
        self._pieces = None

    def add(self, arg, unit=1.0):
        arg = arg / unit
        if self._pieces is None and not _T.is_sym(arg):
            self._condition = np.hstack((self._condition, arg))
            return
        if self._pieces is None:
            self._pieces = [self._condition] if self._condition.size else []
        self._pieces.append(arg)

    def equal(self, arg1, arg2, unit=1.0):
        self.add(arg1 - arg2, unit)

    def lower_bound(self, arg1, arg2, unit=1.0):
        self.add(arg1 - arg2, unit)

    def upper_bound(self, arg1, arg2, unit=1.0):
        self.add(arg2 - arg1, unit)

    def change_value(self, index, value):
        self._condition[index] = value

    def __call__(self):
        if self._pieces is not None:
            return _T.SymRows(self._pieces)
        return self._condition


class Dynamics(object):
    """Per-state right-hand-side holder (reference optimize.py:1075-1127)."""

    def __init__(self, prob, section=0):
        self.section = section
        self.number_of_state = prob.number_of_states[section]
        self.unit_states = prob.unit_states
        self.unit_time = prob.unit_time
        for i in range(self.number_of_state):
            self.__dict__[i] = np.zeros(prob.nodes[section])

    def __getitem__(self, key):
        assert key < self.number_of_state, "Error, Dynamics key out of range"
        return self.__dict__[key]

    def __setitem__(self, key, value):
        assert key < self.number_of_state, "Error, Dynamics key out of range"
        self.__dict__[key] = value

    def __call__(self):
        rhs = [self.__dict__[i] * (self.unit_time / self.unit_states[self.section][i])
               for i in range(self.number_of_state)]
        if any(_T.is_sym(r) for r in rhs):
            return _T.SymDynamics(self.section, rhs)
        dx = np.zeros(0)
        for r in rhs:
            dx = np.hstack((dx, r))
        return dx
