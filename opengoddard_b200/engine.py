"""Device engine: owns the libogb200 problem handle and the PyTorch tensors around it.

PyTorch is plumbing here (device memory, streams, pinned staging buffers); all
arithmetic happens in the sm_100a kernels behind the C ABI (include/ogb200.h).
Everything raises if CUDA or the library is unavailable -- there is no CPU fallback.
"""
import ctypes as C
import os
import warnings

import numpy as np

from . import capi

ABS_STEP = float(np.sqrt(np.finfo(np.float64).eps))     # scipy/optimize/_slsqp_py.py:34


class DeviceProblem:
    """One traced problem resident on one GPU.

    eval(P)    -> c  (B, nrows)                     reference closures, optimize.py:670-715
    eval_fd(P) -> c, J (B, nvars, nrows)            + SciPy's FD Jacobian (_numdiff.py:683-712)
    J[b, j, :] is the column of variable j; rows are [c_eq ; c_ineq ; cost].
    """

    def __init__(self, ir, bounds, device=None, jit=True):
        """jit: compile the traced tapes into the sweep kernel with NVRTC (about 1-2 s once per
        distinct problem; worth it for batches, not for one single-instance solve).
        $OGB200_JIT = 0 forces the interpreter kernel, = require makes a JIT failure an error."""
        import torch
        self.torch = torch
        self.b = capi.ogb()                       # raises OgbError if the .so is missing
        if not torch.cuda.is_available():
            raise capi.OgbError("OpenGoddard-B200 needs a CUDA device (sm_100a); none is "
                                "visible and there is no CPU fallback for the hot path")
        self.device = torch.device(device if device is not None else "cuda:0")
        if self.device.type != "cuda":
            raise capi.OgbError("device must be a CUDA device, not %r" % (self.device,))
        self.ir = ir
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)    # make sure the context exists
            self.h, info = self.b.create(ir)
        self.info = info
        self.nvars, self.meq, self.mineq, self.nrows = info.nvars, info.meq, info.mineq, info.nrows
        self.ndx = info.ndx
        lb, ub = bounds
        assert len(lb) == self.nvars and len(ub) == self.nvars
        self.lb = torch.as_tensor(np.asarray(lb, dtype=np.float64), device=self.device)
        self.ub = torch.as_tensor(np.asarray(ub, dtype=np.float64), device=self.device)
        self._work = None
        self._one = None
        self.fused_dx = -1                        # option 4: -1 automatic (one launch for small batches), 0 K1 + K2, 1 fused
        mode = os.environ.get("OGB200_JIT", "")
        self.jit_error = None
        if mode != "0" and (jit or mode == "require"):
            try:
                self.set_option(2, 1)
                self.b.problem_info_get(self.h, C.byref(self.info))
            except capi.OgbError as e:
                if mode == "require":
                    raise
                self.jit_error = str(e)
                warnings.warn("OpenGoddard-B200: NVRTC specialisation unavailable (%s); using the "
                              "tape-interpreter kernel" % e, RuntimeWarning)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.b.problem_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    @property
    def launches(self):
        """Kernels launched through this handle so far, counted by the library where it launches them."""
        info = capi.OgbProblemInfo()
        self.b.problem_info_get(self.h, C.byref(info))
        return int(info.launches)

    def set_option(self, key, value):
        """ogb_problem_set_option (include/ogb200.h): 0 generic columns, 1 threads, 2 jit,
        3 grid cap, 4 fused D.X, 9 split pipeline (-1 auto / 0 fused / 1 split), 10 split chunk,
        11 streaming zero stores in K2b, 12 zero mode, 13 K1 form (0 latency-organised / 8 round-1 kernel),
        14 programmatic dependent launch of the sweep behind K1, 15 tail refinement (per cent of a wave)."""
        self._rc(self.b.lib.ogb_problem_set_option(self.h, int(key), int(value)), "ogb_problem_set_option")
        if int(key) == 4:
            self.fused_dx = int(value)

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def _workspace(self, B):
        need = self.b.lib.ogb_workspace_bytes(self.h, int(B))
        if self._work is None or self._work.numel() < need:
            self._work = self.torch.empty(need, dtype=self.torch.uint8, device=self.device)
        return self._work

    def _check_P(self, P):
        t = self.torch
        if not isinstance(P, t.Tensor):
            P = t.as_tensor(np.ascontiguousarray(P, dtype=np.float64), device=self.device)
        if P.dim() == 1:
            P = P.unsqueeze(0)
        if P.dtype != t.float64 or P.device != self.device or not P.is_contiguous():
            P = P.to(device=self.device, dtype=t.float64).contiguous()
        if P.shape[1] != self.nvars:
            raise ValueError("P must be (B, %d), got %s" % (self.nvars, tuple(P.shape)))
        return P

    def _rc(self, rc, what):
        if rc != 0:
            raise capi.OgbError("%s failed: %s" % (what, self.b.error()))

    # ------------------------------------------------------------------ device API
    def dx_gemm(self, P, out=None, clip=False):
        """K1 alone: D.X for every phase/state of every instance -> (B, ndx)."""
        t = self.torch
        P = self._check_P(P)
        B = P.shape[0]
        DX = out if out is not None else t.empty((B, self.ndx), dtype=t.float64, device=self.device)
        lb = self.lb.data_ptr() if clip else None
        ub = self.ub.data_ptr() if clip else None
        with t.cuda.device(self.device):
            self._rc(self.b.lib.ogb_dx_gemm(self.h, P.data_ptr(), lb, ub, B, DX.data_ptr(), self._stream()),
                     "ogb_dx_gemm")
        return DX

    def sweep_fd(self, P, DX, out_c, out_J, abs_step=ABS_STEP):
        """K2 alone on a D.X produced by dx_gemm(P, clip=True) (bench / profiling)."""
        t = self.torch
        P = self._check_P(P)
        with t.cuda.device(self.device):
            self._rc(self.b.lib.ogb_sweep(self.h, P.data_ptr(), DX.data_ptr(), self.lb.data_ptr(),
                                          self.ub.data_ptr(), float(abs_step), P.shape[0],
                                          out_c.data_ptr(), out_J.data_ptr(), self._stream()), "ogb_sweep")
        return out_c, out_J

    def eval(self, P, out=None):
        t = self.torch
        P = self._check_P(P)
        B = P.shape[0]
        c = out if out is not None else t.empty((B, self.nrows), dtype=t.float64, device=self.device)
        work = self._workspace(B)
        with t.cuda.device(self.device):
            self._rc(self.b.lib.ogb_eval(self.h, P.data_ptr(), B, c.data_ptr(), work.data_ptr(),
                                         self._stream()), "ogb_eval")
        return c

    def eval_fd(self, P, out_c=None, out_J=None, abs_step=ABS_STEP):
        t = self.torch
        P = self._check_P(P)
        B = P.shape[0]
        c = out_c if out_c is not None else t.empty((B, self.nrows), dtype=t.float64, device=self.device)
        J = out_J if out_J is not None else t.empty((B, self.nvars, self.nrows), dtype=t.float64,
                                                    device=self.device)
        assert c.is_contiguous() and J.is_contiguous()
        work = self._workspace(B)
        with t.cuda.device(self.device):
            self._rc(self.b.lib.ogb_eval_fd(self.h, P.data_ptr(), self.lb.data_ptr(), self.ub.data_ptr(),
                                            float(abs_step), B, c.data_ptr(), J.data_ptr(),
                                            work.data_ptr(), self._stream()), "ogb_eval_fd")
        return c, J

    def eval_sparse(self, P, out_c=None, out_vals=None, abs_step=ABS_STEP):
        """c (B, nrows) and the packed non-zeros of the FD Jacobian, vals (B, nnz) in the jac_pattern()
        layout (ogb_eval_sparse: K1 + the sweep kernel with packed output; no dense J in HBM)."""
        t = self.torch
        P = self._check_P(P)
        B = P.shape[0]
        nnz = self.nnz
        c = out_c if out_c is not None else t.empty((B, self.nrows), dtype=t.float64, device=self.device)
        vals = out_vals if out_vals is not None else t.empty((B, nnz), dtype=t.float64, device=self.device)
        assert c.is_contiguous() and vals.is_contiguous()
        work = self._workspace(B)
        with t.cuda.device(self.device):
            self._rc(self.b.lib.ogb_eval_sparse(self.h, P.data_ptr(), self.lb.data_ptr(), self.ub.data_ptr(),
                                                float(abs_step), B, c.data_ptr(), vals.data_ptr(),
                                                work.data_ptr(), self._stream()), "ogb_eval_sparse")
        return c, vals

    def eval_exact(self, P, out_c=None, out_vals=None):
        """Exact mode (ogb_eval_exact): c (B, nrows) at clip(P) and the structural non-zeros of the exact
        Jacobian -- analytic collocation block, forward-mode tangents of the traced tapes -- packed (B, nnz)
        in the jac_pattern() layout; densify(vals) gives the dense matrix."""
        t = self.torch
        P = self._check_P(P)
        B = P.shape[0]
        c = out_c if out_c is not None else t.empty((B, self.nrows), dtype=t.float64, device=self.device)
        vals = out_vals if out_vals is not None else t.empty((B, self.nnz), dtype=t.float64, device=self.device)
        assert c.is_contiguous() and vals.is_contiguous()
        work = self._workspace(B)
        with t.cuda.device(self.device):
            self._rc(self.b.lib.ogb_eval_exact(self.h, P.data_ptr(), self.lb.data_ptr(), self.ub.data_ptr(), B,
                                               c.data_ptr(), vals.data_ptr(), work.data_ptr(), self._stream()),
                     "ogb_eval_exact")
        return c, vals

    # ------------------------------------------------------------------ guesses / starts / trajectories
    def guess_batch(self, specs, params, time_nodes, tfinal=None, out=None, base=None):
        """ogb_guess_fill: a batch of initial guesses built on the device.  specs: list of
        (kind, block, section) with kind in "zeros" / "constant" / "linear" / "cubic", block = state number or
        nstates + control number, section = phase or None for every phase over the concatenated time axis
        (the reference's Guess.*(prob.time_all_section, ...) + set_*_all_section); params (B, len(specs), 4);
        time_nodes = prob.time_all_section; tfinal (B, nsec) optional.  Entries no spec covers come from
        `base` (n,) (default zeros).  Returns P (B, nvars) on the device."""
        t = self.torch
        params = t.as_tensor(np.ascontiguousarray(params, dtype=np.float64), device=self.device) \
            if not isinstance(params, t.Tensor) else params.to(self.device, t.float64).contiguous()
        B, ns = int(params.shape[0]), len(specs)
        assert tuple(params.shape) == (B, ns, 4)
        tn = t.as_tensor(np.ascontiguousarray(time_nodes, dtype=np.float64), device=self.device)
        assert tn.numel() == self.info.total_nodes
        tf = None
        if tfinal is not None:
            tf = t.as_tensor(np.ascontiguousarray(tfinal, dtype=np.float64), device=self.device) \
                if not isinstance(tfinal, t.Tensor) else tfinal.to(self.device, t.float64).contiguous()
            assert tf.shape[0] == B and tf.numel() == B * len(self.ir.nodes)
        P = out if out is not None else t.zeros((B, self.nvars), dtype=t.float64, device=self.device)
        if base is not None:
            P[:] = t.as_tensor(np.asarray(base, dtype=np.float64), device=self.device)
        arr = (capi.OgbGuessSpec * max(1, ns))()
        for i, (kind, blk, sec) in enumerate(specs):
            arr[i] = capi.OgbGuessSpec(-1 if sec is None else int(sec), int(blk), capi.GUESS_KINDS[kind], 0)
        with t.cuda.device(self.device):
            self._rc(self.b.lib.ogb_guess_fill(self.h, arr, ns, params.data_ptr(), tn.data_ptr(),
                                               tf.data_ptr() if tf is not None else None, B, P.data_ptr(),
                                               self._stream()), "ogb_guess_fill")
        return P

    def jitter_(self, P, seed, first=0, rel_x=0.01, rel_t=0.05, clip=True):
        """ogb_jitter, in place: the seeded multi-start perturbation of a device batch P (B, nvars)."""
        t = self.torch
        assert isinstance(P, t.Tensor) and P.is_contiguous() and P.dtype == t.float64 and P.device == self.device
        with t.cuda.device(self.device):
            self._rc(self.b.lib.ogb_jitter(self.h, P.data_ptr(), int(P.shape[0]), int(seed), int(first), float(rel_x),
                                           float(rel_t), self.lb.data_ptr() if clip else None,
                                           self.ub.data_ptr() if clip else None, self._stream()), "ogb_jitter")
        return P

    def make_starts(self, p0, B, seed, first=0, rel_x=0.01, rel_t=0.05):
        """B perturbed copies of the guess p0 (n,), generated on the device (nothing crosses PCIe but p0)."""
        t = self.torch
        P = t.as_tensor(np.asarray(p0, dtype=np.float64), device=self.device).repeat(int(B), 1).contiguous()
        return self.jitter_(P, seed, first, rel_x, rel_t)

    def trajectories(self, P):
        """ogb_trajectories: (B, total_nodes, 1 + ns + nc) -- time, states, controls per node, dimensional
        (what time_update / states_all_section / controls_all_section / to_csv give per instance)."""
        t = self.torch
        P = self._check_P(P)
        W = 1 + self.ir.nstates[0] + self.ir.ncontrols[0]
        out = t.empty((P.shape[0], self.info.total_nodes, W), dtype=t.float64, device=self.device)
        with t.cuda.device(self.device):
            self._rc(self.b.lib.ogb_trajectories(self.h, P.data_ptr(), int(P.shape[0]), out.data_ptr(), self._stream()),
                     "ogb_trajectories")
        return out

    def densify(self, vals, out_J=None):
        """K2b alone: packed values (B, nnz) -> dense J (B, nvars, nrows), zeros included."""
        t = self.torch
        B = vals.shape[0]
        J = out_J if out_J is not None else t.empty((B, self.nvars, self.nrows), dtype=t.float64, device=self.device)
        assert vals.is_contiguous() and J.is_contiguous() and vals.shape[1] == self.nnz
        with t.cuda.device(self.device):
            self._rc(self.b.lib.ogb_densify(self.h, vals.data_ptr(), B, J.data_ptr(), self._stream()), "ogb_densify")
        return J

    @property
    def nnz(self):
        if getattr(self, "_nnz", None) is None:
            self._nnz = len(self.jac_pattern())
        return self._nnz

    def host_evaluator(self, exact=False):
        """numpy-in / numpy-out evaluator for sqp.slsqp_batch (one per engine and Jacobian flavour: it owns a
        host session).  exact=True: the Jacobians are the exact ones (ogb_eval_exact) instead of FD."""
        cache = self.__dict__.setdefault("_host_evals", {})
        if bool(exact) not in cache:
            cache[bool(exact)] = _HostEvaluator(self, exact=bool(exact))
        return cache[bool(exact)]

    def autotune(self, P, candidates=(256, 384, 128), min_gain=0.03, reps=5):
        """Pick the CTA size of the NVRTC-specialised sweep kernel by timing ogb_sweep on the batch
        P (device or host array): wide CTAs (12 warps) finish the tape phase of problems with heavy
        dynamics in one round instead of two (polar 3 x 40: 0.60 -> 0.56 ms), narrow ones keep more
        registers per thread.  The first candidate stays unless another is faster by `min_gain`.
        Results are bit-identical whatever the choice.  Returns {threads: ms}."""
        t = self.torch
        P = self._check_P(P)
        B = P.shape[0]
        c = t.empty((B, self.nrows), dtype=t.float64, device=self.device)
        J = t.empty((B, self.nvars, self.nrows), dtype=t.float64, device=self.device)
        DX = self.dx_gemm(P, clip=True)
        timings = {}
        for thr in candidates:
            try:
                self.set_option(1, thr)
            except capi.OgbError:
                continue
            for _ in range(2):
                self.sweep_fd(P, DX, c, J)
            best = float("inf")
            for _ in range(reps):
                e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
                e0.record(t.cuda.current_stream(self.device))
                self.sweep_fd(P, DX, c, J)
                e1.record(t.cuda.current_stream(self.device))
                e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            timings[int(thr)] = best
        if not timings:
            raise capi.OgbError("autotune: no candidate CTA size fits this problem")
        first = next(iter(timings))
        choice = first
        for thr, ms in timings.items():
            if ms < timings[choice] and ms < timings[first] * (1.0 - min_gain):
                choice = thr
        self.set_option(1, choice)
        self.b.problem_info_get(self.h, C.byref(self.info))
        self.tuned_threads = choice
        return timings

    def device_sqp(self, max_batch, ftol=1e-6, maxiter=25):
        """The batched SLSQP on the device (ogb_sqp_*): see DeviceSqp."""
        return DeviceSqp(self, max_batch, ftol, maxiter)

    def host_session(self, max_batch, chunk=0, threads=0, exact=False):
        """Host-buffer entry point (ogb_host_eval_fd): numpy / pinned host arrays in and out."""
        return HostSession(self, max_batch, chunk, threads, exact)

    def jac_pattern(self):
        """Ascending linear indices j * nrows + r of the entries of one instance's J that can be
        non-zero (ogb_jac_pattern)."""
        with self.torch.cuda.device(self.device):
            nnz = self.b.lib.ogb_jac_pattern(self.h, None, 0)
            if nnz < 0:
                self._rc(nnz, "ogb_jac_pattern")
            lin = np.empty(nnz, dtype=np.uint32)
            self._rc(min(0, self.b.lib.ogb_jac_pattern(self.h, lin.ctypes.data, nnz)), "ogb_jac_pattern")
        return lin

    def pack(self, J, out=None):
        """K3: dense device J (B, nvars, nrows) -> packed (B, nnz) device tensor."""
        t = self.torch
        B = J.shape[0]
        nnz = self.nnz
        vals = out if out is not None else t.empty((B, nnz), dtype=t.float64, device=self.device)
        with t.cuda.device(self.device):
            self._rc(self.b.lib.ogb_pack(self.h, J.data_ptr(), B, vals.data_ptr(), self._stream()), "ogb_pack")
        return vals

    # ------------------------------------------------------------------ single-instance host API
    # (what Problem.solve hands to SciPy: host vector in, host arrays out, pinned staging)
    def _staging(self):
        if self._one is None:
            t = self.torch
            self._one = dict(
                hx=t.empty((1, self.nvars), dtype=t.float64).pin_memory(),
                hc=t.empty((1, self.nrows), dtype=t.float64).pin_memory(),
                hJ=t.empty((1, self.nvars, self.nrows), dtype=t.float64).pin_memory(),
                dx=t.empty((1, self.nvars), dtype=t.float64, device=self.device),
                dc=t.empty((1, self.nrows), dtype=t.float64, device=self.device),
                dJ=t.empty((1, self.nvars, self.nrows), dtype=t.float64, device=self.device))
        return self._one

    def eval_host(self, x):
        s = self._staging()
        s["hx"].numpy()[0, :] = x
        s["dx"].copy_(s["hx"], non_blocking=True)
        self.eval(s["dx"], out=s["dc"])
        s["hc"].copy_(s["dc"], non_blocking=True)
        self.torch.cuda.current_stream(self.device).synchronize()
        return s["hc"].numpy()[0].copy()

    def eval_fd_host(self, x):
        s = self._staging()
        s["hx"].numpy()[0, :] = x
        s["dx"].copy_(s["hx"], non_blocking=True)
        self.eval_fd(s["dx"], out_c=s["dc"], out_J=s["dJ"])
        s["hc"].copy_(s["dc"], non_blocking=True)
        s["hJ"].copy_(s["dJ"], non_blocking=True)
        self.torch.cuda.current_stream(self.device).synchronize()
        return s["hc"].numpy()[0].copy(), s["hJ"].numpy()[0].copy()


class HostSession:
    """ogb_host_session_* (include/ogb200.h): decision vectors in HOST memory in, c and J in HOST
    memory out, B instances per call, chunked H2D -> K1 -> K2 -> K3 -> D2H (packed) -> host-thread
    expansion.  Arrays are numpy arrays or CPU torch tensors (pinned or pageable), float64,
    C-contiguous.  mode: "dense" (J fully rewritten), "keep_zeros" (J already holds the problem's
    zero background), "packed" (J is (B, nnz)), "dma" (one dense device->host copy)."""

    def __init__(self, eng, max_batch, chunk=0, threads=0, exact=False):
        self.eng = eng
        self.max_batch = int(max_batch)
        self.exact = bool(exact)
        self.lb = np.ascontiguousarray(eng.lb.cpu().numpy())
        self.ub = np.ascontiguousarray(eng.ub.cpu().numpy())
        with eng.torch.cuda.device(eng.device):
            self.h = eng.b.lib.ogb_host_session_create(eng.h, self.max_batch, int(chunk), int(threads))
        if not self.h:
            raise capi.OgbError("ogb_host_session_create failed: " + eng.b.error())
        self.nnz = len(eng.jac_pattern())
        if self.exact:
            eng._rc(eng.b.lib.ogb_host_session_set_option(self.h, 0, 1), "ogb_host_session_set_option")

    def close(self):
        if getattr(self, "h", None):
            self.eng.b.lib.ogb_host_session_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _ptr(a, count):
        if isinstance(a, np.ndarray):
            assert a.dtype == np.float64 and a.flags.c_contiguous and a.size >= count
            return a.ctypes.data
        assert a.device.type == "cpu" and a.is_contiguous() and a.numel() >= count       # torch CPU tensor
        return a.data_ptr()

    def eval_fd(self, P, c=None, J=None, mode="dense", abs_step=ABS_STEP):
        eng = self.eng
        if isinstance(P, np.ndarray):
            P = np.ascontiguousarray(P, dtype=np.float64)
        B = int(P.shape[0])
        n, M = eng.nvars, eng.nrows
        if c is None:
            c = np.empty((B, M), dtype=np.float64)
        if J is None:
            assert mode != "keep_zeros", "keep_zeros needs the caller's prepared J"
            J = np.empty((B, self.nnz) if mode == "packed" else (B, n, M), dtype=np.float64)
        jcount = B * (self.nnz if mode == "packed" else n * M)
        rc = eng.b.lib.ogb_host_eval_fd(self.h, self._ptr(P, B * n), self.lb.ctypes.data, self.ub.ctypes.data,
                                        float(abs_step), B, self._ptr(c, B * M), self._ptr(J, jcount),
                                        capi.HOST_MODES[mode])
        eng._rc(rc, "ogb_host_eval_fd")
        return c, J

    def eval_fd_scatter(self, P, c, C_ptrs, ld, mrows, G_ptrs=None, abs_step=ABS_STEP):
        """ogb_host_eval_fd_scatter: the packed non-zeros of instance b go straight into the caller's
        Fortran-ordered matrix at address C_ptrs[b] (leading dimension ld, rows < mrows) and the cost
        row into the vector at G_ptrs[b] -- the SQP driver's persistent per-instance buffers (which
        keep their zero background: only structural non-zeros are ever written)."""
        eng = self.eng
        P = np.ascontiguousarray(P, dtype=np.float64)
        B = int(P.shape[0])
        n, M = eng.nvars, eng.nrows
        assert len(C_ptrs) == B and (G_ptrs is None or len(G_ptrs) == B)
        cp = (C.c_void_p * max(1, B))(*[int(a) for a in C_ptrs])
        gp = (C.c_void_p * max(1, B))(*[int(a) for a in G_ptrs]) if G_ptrs is not None else None
        rc = eng.b.lib.ogb_host_eval_fd_scatter(self.h, self._ptr(P, B * n), self.lb.ctypes.data, self.ub.ctypes.data,
                                                float(abs_step), B, self._ptr(c, B * M), cp, int(ld), int(mrows), gp)
        eng._rc(rc, "ogb_host_eval_fd_scatter")
        return c

    def stats(self):
        st = capi.OgbHostStats()
        self.eng._rc(self.eng.b.lib.ogb_host_session_stats(self.h, C.byref(st)), "ogb_host_session_stats")
        return st


class _HostEvaluator:
    """numpy in / numpy out view of a DeviceProblem for sqp.slsqp_batch."""

    def __init__(self, eng, exact=False):
        self.eng = eng
        self.exact = exact
        self.session = None

    def eval(self, X):
        t = self.eng.torch
        c = self.eng.eval(t.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).to(self.eng.device))
        return c.cpu().numpy()

    accepts_out = True            # eval_fd(X, out_J=...) writes the Jacobians straight into the caller's buffer

    def eval_fd(self, X, out_J=None):
        """c (k, M) and the dense J (k, n, M) in host memory through the host-buffer session
        (packed device->host transport, ogb_host_eval_fd).  out_J: a C-contiguous (k, n, M) float64
        array (e.g. a slice of the SQP driver's shared-memory block) that receives J."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        k = X.shape[0]
        self._session_for(k)
        return self.session.eval_fd(X, J=out_J, mode="dense")

    def eval_fd_scatter(self, X, C_ptrs, ld, mrows, G_ptrs=None):
        """c (k, M); the Jacobians go straight into the SQP driver's per-instance matrices (see
        HostSession.eval_fd_scatter): what sqp.slsqp_batch uses when the evaluator offers it."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        k = X.shape[0]
        self._session_for(k)
        c = np.empty((k, self.eng.nrows), dtype=np.float64)
        return self.session.eval_fd_scatter(X, c, C_ptrs, ld, mrows, G_ptrs)

    def _session_for(self, k):
        if self.session is None or self.session.max_batch < k:
            if self.session is not None:
                self.session.close()
            self.session = self.eng.host_session(max(k, 16), exact=self.exact)


def scipy_callables(eng, on_x=None, cost_derivative=None, args=()):
    """(fun, constraints, jac) for scipy.optimize.minimize(method='SLSQP') whose values AND Jacobians
    come from the device engine `eng` (anything with eval_host(x) -> c and eval_fd_host(x) -> c, J):
    the replacement of the reference's `for_solver(cost_add / equality_add / inequality)` closures and of
    SciPy's FD `cjac` (reference optimize.py:711-733, scipy/optimize/_slsqp_py.py:353-367).  One device
    evaluation serves every callable asked at the same x (SLSQP asks fun, eq, ineq -- then jac, eq.jac,
    ineq.jac -- at one point).  on_x(x): called with every x SciPy hands over (the facade keeps
    prob.p = x like the reference); cost_derivative(x): optional user gradient of the cost."""
    meq, mineq, M = eng.meq, eng.mineq, eng.nrows
    memo = {"cx": None, "c": None, "jx": None, "jc": None, "J": None}
    on_x = on_x or (lambda x: None)

    def c_at(x):
        on_x(x)
        if memo["jx"] is not None and np.array_equal(memo["jx"], x):
            return memo["jc"]
        if memo["cx"] is None or not np.array_equal(memo["cx"], x):
            memo["c"] = eng.eval_host(x)
            memo["cx"] = np.array(x, copy=True)
        return memo["c"]

    def j_at(x):
        on_x(x)
        if memo["jx"] is None or not np.array_equal(memo["jx"], x):
            memo["jc"], memo["J"] = eng.eval_fd_host(x)
            memo["jx"] = np.array(x, copy=True)
        return memo["J"]

    def fun(x, *a):
        return float(c_at(x)[M - 1])

    cons = ({"type": "eq", "fun": lambda x, *a: c_at(x)[:meq],
             "jac": lambda x, *a: j_at(x)[:, :meq].T, "args": args},
            {"type": "ineq", "fun": lambda x, *a: c_at(x)[meq:meq + mineq],
             "jac": lambda x, *a: j_at(x)[:, meq:meq + mineq].T, "args": args})
    if cost_derivative is None:
        # contiguous copy: SciPy 1.18's low-level SLSQP step reads a strided gradient as if it were
        # contiguous (checked in tests/test_sqp.py), and J[:, M-1] is a strided column
        jac = lambda x, *a: np.ascontiguousarray(j_at(x)[:, M - 1])
    else:
        def jac(x, *a):
            on_x(x)
            return cost_derivative(x)
    return fun, cons, jac


def lgl_device(N, device="cuda:0"):
    """K0 on the device: tau (N,), w (N,), D (N, N) as CUDA tensors."""
    import torch
    b = capi.ogb()
    if not torch.cuda.is_available():
        raise capi.OgbError("ogb_lgl_build needs a CUDA device")
    dev = torch.device(device)
    tau = torch.empty(N, dtype=torch.float64, device=dev)
    w = torch.empty(N, dtype=torch.float64, device=dev)
    D = torch.empty((N, N), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = b.lib.ogb_lgl_build(int(N), tau.data_ptr(), w.data_ptr(), D.data_ptr(),
                                 torch.cuda.current_stream(dev).cuda_stream)
    if rc != 0:
        raise capi.OgbError("ogb_lgl_build failed: " + b.error())
    return tau, w, D


SQP_SCALARS = ("f", "f0", "gs", "h1", "h2", "h3", "h4", "t", "t0", "alpha", "mode", "iter", "reset", "line",
               "inconsistent", "nfev", "njev", "clk_lsei_phase1", "clk_build", "clk_lsei", "clk_lsi", "clk_nnls", "clk_finish", "clk_bfgs")


class SqpKernel:
    """Raw handle of the device SLSQP step (ogb_sqp_create / start / step / scalars) for one problem shape:
    n variables, m constraints (the first meq equalities), the packed Jacobian pattern by variable
    (colptr (n + 1), prow (nnz), rows in [0, m]; row m is the cost gradient) and the bounds."""

    def __init__(self, b, torch, device, n, m, meq, colptr, prow, lb, ub, ftol, maxiter, max_batch):
        self.b, self.torch, self.device = b, torch, device
        self.n, self.m, self.meq, self.max_batch, self.maxiter = int(n), int(m), int(meq), int(max_batch), int(maxiter)
        colptr = np.ascontiguousarray(colptr, dtype=np.int32)
        prow = np.ascontiguousarray(prow, dtype=np.int32)
        lb = np.ascontiguousarray(lb, dtype=np.float64)
        ub = np.ascontiguousarray(ub, dtype=np.float64)
        self.nnz = len(prow)
        with torch.cuda.device(device):
            self.h = b.lib.ogb_sqp_create(self.n, self.m, self.meq, self.nnz, colptr.ctypes.data, prow.ctypes.data,
                                          lb.ctypes.data, ub.ctypes.data, float(ftol), self.maxiter, self.max_batch)
        if not self.h:
            raise capi.OgbError("ogb_sqp_create failed: " + b.error())

    def _rc(self, rc, what):
        if rc != 0:
            raise capi.OgbError("%s failed: %s" % (what, self.b.error()))

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    @property
    def bytes(self):
        return int(self.b.lib.ogb_sqp_bytes(self.h))

    @property
    def launches(self):
        return int(self.b.lib.ogb_sqp_launches(self.h))

    def start(self, B):
        with self.torch.cuda.device(self.device):
            self._rc(self.b.lib.ogb_sqp_start(self.h, int(B), self._stream()), "ogb_sqp_start")

    def step(self, X, c, vals, mode):
        """X (B, n) device in / out; c (B, m + 1), vals (B, nnz) device; mode (B,) int32 host array out."""
        t = self.torch
        B = X.shape[0]
        assert X.is_contiguous() and c.is_contiguous() and vals.is_contiguous() and X.dtype == t.float64
        assert tuple(c.shape) == (B, self.m + 1) and tuple(vals.shape) == (B, self.nnz) and mode.dtype == np.int32
        with t.cuda.device(self.device):
            self._rc(self.b.lib.ogb_sqp_step(self.h, X.data_ptr(), c.data_ptr(), vals.data_ptr(), B, mode.ctypes.data,
                                             self._stream()), "ogb_sqp_step")

    def scalars(self, B):
        sc = np.empty((B, 24), dtype=np.float64)
        with self.torch.cuda.device(self.device):
            self._rc(self.b.lib.ogb_sqp_scalars(self.h, int(B), sc.ctypes.data, self._stream()), "ogb_sqp_scalars")
        return {name: sc[:, i] for i, name in enumerate(SQP_SCALARS)}

    def close(self):
        if getattr(self, "h", None):
            with self.torch.cuda.device(self.device):
                self.b.lib.ogb_sqp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceSqp:
    """B independent SLSQP runs advanced in lock step entirely on the GPU (SURVEY.md section 8f row 1,
    "device-side batched QP"): every round is one packed sweep (ogb_eval_sparse / ogb_eval_exact: c and the
    Jacobian of every instance at its current x) and one ogb_sqp_step launch in which a thread block per
    instance runs Kraft's SLSQP step -- BFGS update, the QP by LSQ/LSEI/LSI/LDP/NNLS, line search, convergence
    tests (csrc/ogb_sqp.h) -- on those device buffers; only one int per instance returns to the host.
    Opt-in: the default `solve_batch` keeps SciPy's compiled core per instance (bitwise SciPy)."""

    def __init__(self, eng, max_batch, ftol=1e-6, maxiter=25):
        self.eng, self.torch = eng, eng.torch
        M, n = eng.nrows, eng.nvars
        if n > 512:
            warnings.warn("OpenGoddard-B200: the device SLSQP keeps one instance's QP matrices per thread block in L2; at "
                          "%d variables a QP streams them from HBM and is no faster than the host SLSQP pool "
                          "(solve_batch's default)" % n, RuntimeWarning, stacklevel=3)
        lin = eng.jac_pattern().astype(np.int64)
        colptr = np.searchsorted(lin, np.arange(n + 1, dtype=np.int64) * M)
        self.k = SqpKernel(eng.b, eng.torch, eng.device, n, M - 1, eng.meq, colptr, lin % M, eng.lb.cpu().numpy(),
                           eng.ub.cpu().numpy(), ftol, maxiter, max_batch)
        self.max_batch, self.maxiter = int(max_batch), int(maxiter)

    @property
    def bytes(self):
        return self.k.bytes

    @property
    def launches(self):
        return self.k.launches

    def solve(self, X0, exact=False, max_rounds=None, callback=None):
        """Run every row of X0 (B, nvars; host or device) to SLSQP's exit.  Returns dict(x (B, n) numpy, fun,
        status, nit, nfev, njev, rounds)."""
        t, eng = self.torch, self.eng
        X = eng._check_P(X0).clone()
        B = X.shape[0]
        assert 0 < B <= self.max_batch
        X = t.minimum(t.maximum(X, eng.lb), eng.ub).contiguous()     # scipy/optimize/_slsqp_py.py:322
        c = t.empty((B, eng.nrows), dtype=t.float64, device=eng.device)
        vals = t.empty((B, eng.nnz), dtype=t.float64, device=eng.device)
        mode = np.zeros(B, dtype=np.int32)
        evaluate = eng.eval_exact if exact else eng.eval_sparse
        rounds = 0
        # every SLSQP iteration is one gradient round and at least one line-search round (at most 11)
        cap = max_rounds or (self.maxiter + 6) * 13
        self.k.start(B)
        while True:
            evaluate(X, out_c=c, out_vals=vals)
            self.k.step(X, c, vals, mode)
            rounds += 1
            if callback is not None:
                callback(rounds, mode)
            if not (np.abs(mode) == 1).any() or rounds >= cap:
                break
        sc = self.k.scalars(B)
        fun = c[:, eng.nrows - 1].cpu().numpy()                  # (c holds the values at the final x of every instance)
        return {"x": X.cpu().numpy(), "fun": fun, "status": sc["mode"].astype(int), "nit": sc["iter"].astype(int),
                "nfev": sc["nfev"].astype(int), "njev": sc["njev"].astype(int), "rounds": rounds}

    def close(self):
        self.k.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
