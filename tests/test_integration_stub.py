"""INTEGRATION.md shows the ctypes stubs a maintainer of the reference would add.  The host-buffer
stub is extracted from the document and executed here, so the documentation cannot drift from the
library: CPU part = it parses and binds every symbol it names; GPU part = it returns what the
engine returns."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from opengoddard_b200 import capi
from tests.helpers import ROOT


def stub_source(marker):
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    for b in blocks:
        if marker in b:
            return b
    raise AssertionError("no python block with %r in INTEGRATION.md" % marker)


def test_stubs_parse_and_name_exported_symbols():
    lib = C.CDLL(capi.LIB_PATH)
    for marker in ("ogb200_stub.py", "ogb200_host_stub.py"):
        src = stub_source(marker)
        compile(src, marker, "exec")
        for name in set(re.findall(r"lib\.(ogb_[a-z0-9_]+)", src)):
            assert hasattr(lib, name), "%s names %s, which libogb200.so does not export" % (marker, name)


@pytest.mark.gpu
def test_host_stub_from_the_document_runs(api):
    from opengoddard_b200 import tape, workloads
    src = stub_source("ogb200_host_stub.py").replace('C.CDLL("libogb200.so")', "C.CDLL(%r)" % capi.LIB_PATH)
    ns = {}
    exec(compile(src, "ogb200_host_stub.py", "exec"), ns)
    wl = workloads.build("cfg2_goddard50", api)
    eng = wl.prob.compile(wl.obj)                         # the engine's own answer
    P = workloads.make_batch(wl, 11)
    c_ref, J_ref = eng.eval_fd(P)
    desc, keep = capi.make_desc(tape.build_ir(wl.prob, wl.obj))
    lb, ub = wl.prob.bounds_arrays()
    hj = ns["HostJacobian"](desc, lb, ub, eng.nvars, eng.nrows, 16)
    c, J = hj(P)
    assert (c == c_ref.cpu().numpy()).all() and (J == J_ref.cpu().numpy()).all()
