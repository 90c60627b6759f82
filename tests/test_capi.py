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
    assert ctypes.sizeof(capi.OgbProblemInfo) == 56          # 12 x int32 + int64 launches
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


def test_product_never_imports_the_oracle_or_the_reference():
    """oracle/ is test infrastructure and /root/reference does not exist on the GPU box: neither may
    be referenced by the product packages, bench.py's CUDA arm or the C sources."""
    import ast
    import glob
    pkg = glob.glob(os.path.join(ROOT, "opengoddard_b200", "*.py")) + glob.glob(os.path.join(ROOT, "OpenGoddard", "*.py"))
    assert len(pkg) >= 10
    for path in pkg:
        tree = ast.parse(open(path).read())
        docstrings = set()                                 # citations of reference file:line live in docstrings
        for node in ast.walk(tree):
            if isinstance(node, (ast.Module, ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef)) and node.body and \
                    isinstance(node.body[0], ast.Expr) and isinstance(node.body[0].value, ast.Constant):
                docstrings.add(id(node.body[0].value))
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n.split(".")[0] in ("oracle", "tests") for n in names), "%s imports %s" % (path, names)
            if isinstance(node, ast.Constant) and isinstance(node.value, str) and id(node) not in docstrings:
                assert "/root/reference" not in node.value, "%s uses the reference tree at run time" % path
    for path in glob.glob(os.path.join(ROOT, "opengoddard_b200", "csrc", "*")) + [HEADER]:
        if path.endswith(".inc"):
            continue
        text = open(path).read()
        assert "#include \"../../oracle" not in text and "oracle/" not in text.replace("// oracle", "")
    # in a fresh interpreter, importing the facade and the engine does not pull the oracle in
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import OpenGoddard.optimize, opengoddard_b200.engine, "
            "opengoddard_b200.sqp, opengoddard_b200.batch; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True, timeout=300)


def test_jit_disk_cache_is_opt_in_and_hits(api, tmp_path, monkeypatch):
    """$OGB200_JIT_CACHE: the NVRTC result is stored under a key made of the generated source, the
    embedded headers, the kernel variant and the library version; a second compile is a file read."""
    import time
    from opengoddard_b200 import tape, workloads
    wl = workloads.build("cfg3_goddard_knot30x2", api)
    ir = tape.build_ir(wl.prob, wl.obj)
    try:
        n0, _ = capi.jit_check(ir)                       # no cache configured: nothing is written anywhere
    except capi.OgbError as e:
        if "libnvrtc not found" in str(e):
            pytest.skip("NVRTC is not installed here")
        raise
    assert list(tmp_path.iterdir()) == []
    monkeypatch.setenv("OGB200_JIT_CACHE", str(tmp_path))
    n1, _ = capi.jit_check(ir)
    files = list(tmp_path.iterdir())
    assert len(files) == 1 and files[0].name.startswith("ogb200_") and files[0].name.endswith(".cubin")
    t0 = time.perf_counter()
    n2, _ = capi.jit_check(ir)
    assert time.perf_counter() - t0 < 0.5 and n0 == n1 == n2
    other = workloads.build("cfg2_goddard50", api)       # a different problem gets a different entry
    capi.jit_check(tape.build_ir(other.prob, other.obj))
    assert len(list(tmp_path.iterdir())) == 2
