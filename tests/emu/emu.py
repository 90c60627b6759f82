"""ctypes access to tests/emu/libogb_emu.so -- TEST INFRASTRUCTURE ONLY (see ogb_emu.cpp)."""
import ctypes as C
import os

import numpy as np

from opengoddard_b200 import capi

EMU_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libogb_emu.so")
_b = None


def binding():
    global _b
    if _b is None:
        from opengoddard_b200 import build
        build.build_emu()
        b = capi.Binding(EMU_LIB, "emu_")
        dp = C.POINTER(C.c_double)
        b.lib.emu_eval.restype = C.c_int
        b.lib.emu_eval.argtypes = [C.c_void_p, dp, dp, dp, C.c_double, C.c_int, dp, dp, C.c_int]
        b.lib.emu_eval_exact.restype = C.c_int
        b.lib.emu_eval_exact.argtypes = [C.c_void_p, dp, dp, dp, C.c_int, dp, dp]
        b.lib.emu_guess_value.restype = C.c_double
        b.lib.emu_guess_value.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, dp]
        b.lib.emu_philox4x32.restype = None
        b.lib.emu_philox4x32.argtypes = [C.POINTER(C.c_uint32)] * 3
        b.lib.emu_u53.restype = C.c_double
        b.lib.emu_u53.argtypes = [C.c_uint32, C.c_uint32]
        b.lib.emu_dx_gemm.restype = C.c_int
        b.lib.emu_dx_gemm.argtypes = [C.c_void_p, dp, C.c_int, dp]
        b.lib.emu_lgl_build.restype = C.c_int
        b.lib.emu_lgl_build.argtypes = [C.c_int, dp, dp, dp]
        _b = b
    return _b


def _ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class EmuProblem:
    def __init__(self, ir, lb, ub):
        self.b = binding()
        self.h, self.info = self.b.create(ir)
        self.lb = np.ascontiguousarray(lb, dtype=np.float64)
        self.ub = np.ascontiguousarray(ub, dtype=np.float64)

    def eval_fd(self, P, abs_step=1.4901161193847656e-08):
        P = np.ascontiguousarray(np.atleast_2d(P), dtype=np.float64)
        B, n, M = P.shape[0], self.info.nvars, self.info.nrows
        assert P.shape[1] == n
        c = np.empty((B, M))
        J = np.empty((B, n, M))
        rc = self.b.lib.emu_eval(self.h, _ptr(P), _ptr(self.lb), _ptr(self.ub), abs_step, B,
                                 _ptr(c), _ptr(J), 1)
        assert rc == 0
        return c, J

    def eval_exact(self, P):
        """c at clip(P) and the exact (forward-mode / analytic) Jacobian, dense (B, n, M)."""
        P = np.ascontiguousarray(np.atleast_2d(P), dtype=np.float64)
        B, n, M = P.shape[0], self.info.nvars, self.info.nrows
        c = np.empty((B, M))
        J = np.empty((B, n, M))
        rc = self.b.lib.emu_eval_exact(self.h, _ptr(P), _ptr(self.lb), _ptr(self.ub), B, _ptr(c), _ptr(J))
        assert rc == 0
        return c, J

    def eval(self, P):
        P = np.ascontiguousarray(np.atleast_2d(P), dtype=np.float64)
        B, M = P.shape[0], self.info.nrows
        c = np.empty((B, M))
        rc = self.b.lib.emu_eval(self.h, _ptr(P), _ptr(self.lb), _ptr(self.ub), 0.0, B,
                                 _ptr(c), _ptr(c), 0)
        assert rc == 0
        return c

    def __del__(self):
        try:
            self.b.problem_destroy(self.h)
        except Exception:
            pass


class EmuSqp:
    """The device SQP (csrc/ogb_sqp.h) run by one serial thread: B instances of one problem shape."""
    KEYS = ("x0", "s", "g", "mu", "r", "v", "u", "w", "lt", "dg", "sc", "total")

    def __init__(self, n, m, meq, colptr, prow, xl, xu, acc, itermax, B):
        b = binding()
        L = b.lib
        ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
        L.emu_sqp_create.restype = C.c_void_p
        L.emu_sqp_create.argtypes = [C.c_int] * 4 + [ip, ip, dp, dp, C.c_double, C.c_int, C.c_int]
        L.emu_sqp_destroy.argtypes = [C.c_void_p]
        L.emu_sqp_offsets.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        L.emu_sqp_state.restype = dp
        L.emu_sqp_state.argtypes = [C.c_void_p, C.c_int]
        L.emu_sqp_step.argtypes = [C.c_void_p, dp, dp, dp]
        L.emu_sqp_lsq.restype = C.c_int
        L.emu_sqp_lsq.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, C.c_int, C.c_double]
        self.L, self.n, self.m, self.meq, self.B = L, n, m, meq, B
        self.colptr = np.ascontiguousarray(colptr, dtype=np.int32)
        self.prow = np.ascontiguousarray(prow, dtype=np.int32)
        self.nnz = len(self.prow)
        xl = np.ascontiguousarray(xl, dtype=np.float64)
        xu = np.ascontiguousarray(xu, dtype=np.float64)
        self.h = L.emu_sqp_create(n, m, meq, self.nnz, self.colptr.ctypes.data_as(ip), self.prow.ctypes.data_as(ip),
                                  _ptr(xl), _ptr(xu), acc, itermax, B)
        if not self.h:
            raise RuntimeError(b.error())
        off = (C.c_longlong * 12)()
        L.emu_sqp_offsets(self.h, off)
        self.off = dict(zip(self.KEYS, [int(v) for v in off]))

    def state(self, b):
        p = self.L.emu_sqp_state(self.h, b)
        return np.ctypeslib.as_array(p, shape=(self.off["total"],))

    def field(self, b, key, count):
        return self.state(b)[self.off[key]:self.off[key] + count]

    def scalars(self, b):
        sc = self.field(b, "sc", 24)
        names = ("f", "f0", "gs", "h1", "h2", "h3", "h4", "t", "t0", "alpha", "mode", "iter", "reset", "line", "badlin",
                 "nfev", "njev")
        return {k: sc[i] for i, k in enumerate(names)}

    def step(self, X, c, vals):
        assert X.flags.c_contiguous and c.flags.c_contiguous and vals.flags.c_contiguous
        self.L.emu_sqp_step(self.h, _ptr(X), _ptr(c), _ptr(vals))

    def lsq(self, b, x, c, vals, aug=False, rho=100.0):
        return self.L.emu_sqp_lsq(self.h, b, _ptr(x), _ptr(c), _ptr(vals), int(aug), float(rho))

    def __del__(self):
        try:
            if self.h:
                self.L.emu_sqp_destroy(self.h)
        except Exception:
            pass
