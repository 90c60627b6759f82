"""ctypes binding of libogb200.so (include/ogb200.h).  No torch types cross this
boundary: device buffers are passed as raw addresses (`tensor.data_ptr()`), the
stream as the raw `cudaStream_t`.

The library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a).  If it is
missing the import of anything that needs it raises -- there is no fallback.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libogb200.so")


class OgbOut(C.Structure):
    _fields_ = [("kind", C.c_int32), ("row", C.c_int32), ("glo", C.c_int32), ("ghi", C.c_int32)]


class OgbProgram(C.Structure):
    _fields_ = [("code_h", C.POINTER(C.c_uint64)), ("ncode", C.c_int32),
                ("consts_h", C.POINTER(C.c_double)), ("nconsts", C.c_int32),
                ("outs_h", C.POINTER(OgbOut)), ("nouts", C.c_int32),
                ("nreg", C.c_int32),
                ("nodec_h", C.POINTER(C.c_double)), ("n_nodec", C.c_int32),
                ("globals_h", C.POINTER(C.c_int32)), ("nglobals", C.c_int32)]


class OgbTable(C.Structure):
    _fields_ = [("off", C.c_int32), ("len", C.c_int32), ("variant", C.c_int32), ("extrapolate", C.c_int32),
                ("fill_below", C.c_double), ("fill_above", C.c_double)]


class OgbGuessSpec(C.Structure):
    _fields_ = [("sec", C.c_int32), ("blk", C.c_int32), ("kind", C.c_int32), ("pad", C.c_int32)]


GUESS_KINDS = {"zeros": 0, "constant": 1, "linear": 2, "cubic": 3}       # enum ogb_guess_kind


class OgbProblemDesc(C.Structure):
    _fields_ = [("nsec", C.c_int32),
                ("nodes_h", C.POINTER(C.c_int32)),
                ("nstates_h", C.POINTER(C.c_int32)),
                ("ncontrols_h", C.POINTER(C.c_int32)),
                ("unit_states_h", C.POINTER(C.c_double)),
                ("unit_time", C.c_double),
                ("t0", C.c_double),
                ("knot_smooth_h", C.POINTER(C.c_uint8)),
                ("meq_user", C.c_int32),
                ("mineq_user", C.c_int32),
                ("has_running_cost", C.c_int32),
                ("node_prog_h", C.POINTER(OgbProgram)),
                ("scalar_prog_h", C.POINTER(OgbProgram)),
                ("ntables", C.c_int32),
                ("tables_h", C.POINTER(OgbTable)),
                ("table_x_h", C.POINTER(C.c_double)),
                ("table_y_h", C.POINTER(C.c_double)),
                ("unit_controls_h", C.POINTER(C.c_double))]


class OgbProblemInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("nvars", "meq", "mineq", "nrows", "ndx", "total_nodes",
                                          "tile_cols", "group_cols", "smem_bytes", "ctas_per_sm", "jit",
                                          "nnz")] + [("launches", C.c_int64)]


class OgbHostStats(C.Structure):
    _fields_ = [("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("launches", C.c_int32),
                ("nnz", C.c_int32), ("chunk", C.c_int32), ("threads", C.c_int32), ("nchunks", C.c_int32),
                ("pad", C.c_int32), ("ms_total", C.c_double), ("ms_first_chunk", C.c_double)]


HOST_MODES = {"dense": 0, "keep_zeros": 1, "packed": 2, "dma": 3}     # enum ogb_host_mode


class OgbError(RuntimeError):
    pass


def _program(tape, keep):
    code = np.ascontiguousarray(tape.code, dtype=np.uint64)
    consts = np.ascontiguousarray(tape.consts if len(tape.consts) else np.zeros(1), dtype=np.float64)
    outs = (OgbOut * max(1, len(tape.outs)))()
    for i, (kind, row, glo, ghi) in enumerate(tape.outs):
        outs[i] = OgbOut(kind, row, glo, ghi)
    nodec = getattr(tape, "nodec", None) or []
    gl = getattr(tape, "globals", None) or []
    ncv = np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=np.float64) for v in nodec]) if nodec else np.zeros(1))
    glv = np.ascontiguousarray(gl if gl else [0], dtype=np.int32)
    keep.extend([code, consts, outs, ncv, glv])
    return OgbProgram(code.ctypes.data_as(C.POINTER(C.c_uint64)), len(code),
                      consts.ctypes.data_as(C.POINTER(C.c_double)), len(tape.consts),
                      outs, len(tape.outs), tape.nreg,
                      ncv.ctypes.data_as(C.POINTER(C.c_double)), len(nodec),
                      glv.ctypes.data_as(C.POINTER(C.c_int32)), len(gl))


def make_desc(ir):
    """ProblemIR (tape.py) -> (OgbProblemDesc, keepalive list)."""
    keep = []

    def arr(values, ctype, dtype):
        a = np.ascontiguousarray(values if len(values) else [0], dtype=dtype)
        keep.append(a)
        return a.ctypes.data_as(C.POINTER(ctype))

    progs = (OgbProgram * len(ir.node_tapes))(*[_program(t, keep) for t in ir.node_tapes])
    scalar = (OgbProgram * 1)(_program(ir.scalar_tape, keep))
    keep.extend([progs, scalar])
    tabs = (OgbTable * max(1, len(ir.tables)))()
    tx, ty, off = [], [], 0
    for i, t in enumerate(ir.tables):
        tabs[i] = OgbTable(off, len(t["x"]), int(t["variant"]), 1 if t["extrapolate"] else 0,
                           float(t["fill_below"]), float(t["fill_above"]))
        tx.extend(np.asarray(t["x"], dtype=float).tolist())
        ty.extend(np.asarray(t["y"], dtype=float).tolist())
        off += len(t["x"])
    keep.append(tabs)
    desc = OgbProblemDesc(
        len(ir.nodes), arr(ir.nodes, C.c_int32, np.int32), arr(ir.nstates, C.c_int32, np.int32),
        arr(ir.ncontrols, C.c_int32, np.int32), arr(ir.unit_states, C.c_double, np.float64),
        float(ir.unit_time), float(ir.t0),
        arr([1 if k else 0 for k in ir.knot_smooth], C.c_uint8, np.uint8),
        int(ir.meq_user), int(ir.mineq_user), 1 if ir.has_running_cost else 0, progs, scalar,
        len(ir.tables), tabs, arr(tx, C.c_double, np.float64), arr(ty, C.c_double, np.float64),
        arr(getattr(ir, "unit_controls", None) or [1.0] * max(1, sum(ir.ncontrols)), C.c_double, np.float64))
    return desc, keep


class Binding:
    """Typed access to the entry points of a library exporting `<prefix>*` symbols."""

    def __init__(self, path, prefix="ogb_"):
        if not os.path.isfile(path):
            raise OgbError("%s not found -- build it with `python __graft_entry__.py build` "
                           "(nvcc, sm_100a); OpenGoddard-B200 has no CPU fallback" % path)
        self.path, self.prefix = path, prefix
        self.lib = C.CDLL(path)
        f = lambda n: getattr(self.lib, prefix + n)
        self.last_error = f("last_error")
        self.last_error.restype = C.c_char_p
        self.problem_create = f("problem_create")
        self.problem_create.restype = C.c_void_p
        self.problem_create.argtypes = [C.POINTER(OgbProblemDesc)]
        self.problem_destroy = f("problem_destroy")
        self.problem_destroy.restype = None
        self.problem_destroy.argtypes = [C.c_void_p]
        self.problem_info_get = f("problem_info_get")
        self.problem_info_get.restype = C.c_int
        self.problem_info_get.argtypes = [C.c_void_p, C.POINTER(OgbProblemInfo)]

    def error(self):
        msg = self.last_error()
        return msg.decode() if msg else ""

    def create(self, ir):
        desc, keep = make_desc(ir)
        h = self.problem_create(C.byref(desc))
        if not h:
            raise OgbError("ogb_problem_create failed: " + self.error())
        info = OgbProblemInfo()
        self.problem_info_get(h, C.byref(info))
        return h, info


_ogb = None


def ogb():
    """The product library (singleton).  Raises OgbError if it has not been built."""
    global _ogb
    if _ogb is None:
        b = Binding(LIB_PATH, "ogb_")
        L = b.lib
        vp, dp, i32 = C.c_void_p, C.c_void_p, C.c_int
        L.ogb_version.restype = C.c_int
        L.ogb_lgl_build.restype = C.c_int
        L.ogb_lgl_build.argtypes = [i32, dp, dp, dp, vp]
        L.ogb_lgl_build_host.restype = C.c_int
        L.ogb_lgl_build_host.argtypes = [i32, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                         C.POINTER(C.c_double)]
        L.ogb_problem_set_option.restype = C.c_int
        L.ogb_problem_set_option.argtypes = [vp, i32, i32]
        L.ogb_jit_check.restype = C.c_int
        L.ogb_jit_check.argtypes = [C.POINTER(OgbProblemDesc), C.c_char_p, C.c_int]
        L.ogb_jit_check_variant.restype = C.c_int
        L.ogb_jit_check_variant.argtypes = [C.POINTER(OgbProblemDesc), C.c_int, C.c_char_p, C.c_int]
        L.ogb_eval_exact.restype = C.c_int
        L.ogb_eval_exact.argtypes = [vp, dp, dp, dp, i32, dp, dp, vp, vp]
        L.ogb_workspace_bytes.restype = C.c_size_t
        L.ogb_workspace_bytes.argtypes = [vp, i32]
        L.ogb_dx_gemm.restype = C.c_int
        L.ogb_dx_gemm.argtypes = [vp, dp, dp, dp, i32, dp, vp]
        L.ogb_sweep.restype = C.c_int
        L.ogb_sweep.argtypes = [vp, dp, dp, dp, dp, C.c_double, i32, dp, dp, vp]
        L.ogb_eval.restype = C.c_int
        L.ogb_eval.argtypes = [vp, dp, i32, dp, vp, vp]
        L.ogb_eval_fd.restype = C.c_int
        L.ogb_eval_fd.argtypes = [vp, dp, dp, dp, C.c_double, i32, dp, dp, vp, vp]
        L.ogb_eval_sparse.restype = C.c_int
        L.ogb_eval_sparse.argtypes = [vp, dp, dp, dp, C.c_double, i32, dp, dp, vp, vp]
        L.ogb_densify.restype = C.c_int
        L.ogb_densify.argtypes = [vp, dp, i32, dp, vp]
        L.ogb_host_eval_fd_scatter.restype = C.c_int
        L.ogb_host_eval_fd_scatter.argtypes = [vp, vp, vp, vp, C.c_double, i32, vp, vp, i32, i32, vp]
        L.ogb_guess_fill.restype = C.c_int
        L.ogb_guess_fill.argtypes = [vp, C.POINTER(OgbGuessSpec), i32, dp, dp, dp, i32, dp, vp]
        L.ogb_jitter.restype = C.c_int
        L.ogb_jitter.argtypes = [vp, dp, i32, C.c_uint64, C.c_int64, C.c_double, C.c_double, dp, dp, vp]
        L.ogb_trajectories.restype = C.c_int
        L.ogb_trajectories.argtypes = [vp, dp, i32, dp, vp]
        L.ogb_jac_pattern.restype = C.c_int
        L.ogb_jac_pattern.argtypes = [vp, vp, i32]
        L.ogb_pack.restype = C.c_int
        L.ogb_pack.argtypes = [vp, dp, i32, dp, vp]
        L.ogb_host_session_create.restype = C.c_void_p
        L.ogb_host_session_create.argtypes = [vp, i32, i32, i32]
        L.ogb_host_session_destroy.restype = None
        L.ogb_host_session_destroy.argtypes = [vp]
        L.ogb_host_eval_fd.restype = C.c_int
        L.ogb_host_eval_fd.argtypes = [vp, vp, vp, vp, C.c_double, i32, vp, vp, i32]
        L.ogb_host_session_set_option.restype = C.c_int
        L.ogb_host_session_set_option.argtypes = [vp, i32, i32]
        L.ogb_host_session_stats.restype = C.c_int
        L.ogb_host_session_stats.argtypes = [vp, C.POINTER(OgbHostStats)]
        L.ogb_host_expand.restype = C.c_int
        L.ogb_host_expand.argtypes = [vp, vp, i32, C.c_size_t, i32, vp, i32, i32]
        L.ogb_sqp_create.restype = C.c_void_p
        L.ogb_sqp_create.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp, C.c_double, i32, i32]
        L.ogb_sqp_destroy.restype = None
        L.ogb_sqp_destroy.argtypes = [vp]
        L.ogb_sqp_bytes.restype = C.c_size_t
        L.ogb_sqp_bytes.argtypes = [vp]
        L.ogb_sqp_start.restype = C.c_int
        L.ogb_sqp_start.argtypes = [vp, i32, vp]
        L.ogb_sqp_step.restype = C.c_int
        L.ogb_sqp_step.argtypes = [vp, vp, vp, vp, i32, vp, vp]
        L.ogb_sqp_scalars.restype = C.c_int
        L.ogb_sqp_scalars.argtypes = [vp, i32, vp, vp]
        L.ogb_sqp_launches.restype = C.c_longlong
        L.ogb_sqp_launches.argtypes = [vp]
        _ogb = b
    return _ogb


def host_expand(vals, lin, nM, out=None, mode="dense", threads=0):
    """ogb_host_expand: packed values (B, nnz) + ascending pattern `lin` -> dense (B, nM) host array
    (the host half of the packed Jacobian transport; runs without a GPU)."""
    b = ogb()
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    lin = np.ascontiguousarray(lin, dtype=np.uint32)
    B, nnz = vals.shape
    assert lin.shape == (nnz,)
    if out is None:
        assert mode == "dense"
        out = np.empty((B, int(nM)), dtype=np.float64)
    assert out.flags.c_contiguous and out.dtype == np.float64 and out.size == B * int(nM)
    rc = b.lib.ogb_host_expand(vals.ctypes.data, lin.ctypes.data, nnz, int(nM), B, out.ctypes.data,
                               HOST_MODES[mode], int(threads))
    if rc != 0:
        raise OgbError("ogb_host_expand failed: " + b.error())
    return out


def lgl_host(N):
    """tau, w, D of the N-point LGL rule, computed by libogb200's host entry point (the
    same __host__ __device__ code the device kernel runs)."""
    b = ogb()
    tau = np.empty(N)
    w = np.empty(N)
    D = np.empty((N, N))
    ptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rc = b.lib.ogb_lgl_build_host(int(N), ptr(tau), ptr(w), ptr(D))
    if rc != 0:
        raise OgbError("ogb_lgl_build_host failed: " + b.error())
    return tau, w, D


def jit_check(ir, variant=0):
    """NVRTC-compile the specialised sweep kernel for a traced problem (no GPU needed); variant 0 = dense /
    c only, 1 = packed FD output, 2 = exact mode.  Returns (cubin_bytes, generated_source); raises OgbError
    with the compiler log."""
    b = ogb()
    desc, keep = make_desc(ir)
    buf = C.create_string_buffer(1 << 21)
    rc = b.lib.ogb_jit_check_variant(C.byref(desc), int(variant), buf, len(buf))
    if rc <= 0:
        raise OgbError("ogb_jit_check failed: " + b.error())
    return rc, buf.value.decode()
