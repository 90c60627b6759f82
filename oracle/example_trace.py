"""Run the reference's shipped example scripts UNCHANGED against the drop-in facade and record the
traced problem IR of each as a fixture (tests/golden/example_XX_ir.npz).

TEST INFRASTRUCTURE ONLY; needs /root/reference/examples (the build container):
    python -m oracle.example_trace

`/root/reference` does not exist on the GPU box, so the `-m gpu` tests cannot exec() the scripts
there.  What they can do is rebuild the device engine from the IR our tracer extracted from the
unchanged script (generated data: opcodes and constants, no reference source), run it through the
CUDA kernels and compare with what the REFERENCE computed for the same script
(tests/golden/example_XX.npz, oracle/make_golden.py).
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

EXDIR = os.path.join(ref_loader.REFERENCE_ROOT, "examples")
# where the scripts can be RUN (they write figures / CSV files next to themselves): the copy that sits beside the
# pip-installed reference under baseline/_ref (git-ignored, travels to the GPU box), else the reference tree
_RUNDIR = os.path.join(ROOT, "baseline", "_ref", "examples")
RUNDIR = _RUNDIR if os.path.isdir(_RUNDIR) else EXDIR
GOLD = os.path.join(ROOT, "tests", "golden")
TAGS = ("01", "02", "03", "04", "05", "06", "07", "08", "09", "10", "11")


def run_script(tag, intercept=True, env_backend=None, exdir=None):
    """exec() one shipped example with `from OpenGoddard.optimize import ...` resolving to the facade.
    intercept: Problem.solve only records (prob, obj, options).  Returns (box, globals, stdout)."""
    import OpenGoddard.optimize as api
    ref_loader.install_matplotlib_stub()
    ref_loader.install_scipy_shims()
    exdir = exdir or EXDIR
    script = [f for f in sorted(os.listdir(exdir)) if f.startswith(tag) and f.endswith(".py")][0]
    box = {}
    real_solve = api.Problem.solve

    def fake_solve(self, obj, display_func=None, **options):
        box["prob"], box["obj"], box["options"] = self, obj, options

    cwd = os.getcwd()
    os.chdir(exdir)
    if intercept:
        api.Problem.solve = fake_solve
    old = os.environ.get("OGB200_BACKEND")
    if env_backend:
        os.environ["OGB200_BACKEND"] = env_backend
    out = io.StringIO()
    try:
        glb = {"__name__": "__main__", "__file__": script}
        with contextlib.redirect_stdout(out):
            try:
                exec(compile(open(script).read(), script, "exec"), glb)
            except Exception:
                if intercept and "prob" in box:
                    pass                      # post-processing on an unsolved problem may fail
                else:
                    raise
    finally:
        os.chdir(cwd)
        api.Problem.solve = real_solve
        if env_backend:
            if old is None:
                os.environ.pop("OGB200_BACKEND", None)
            else:
                os.environ["OGB200_BACKEND"] = old
    return box, glb, out.getvalue()


def main():
    from opengoddard_b200 import tape
    for tag in TAGS:
        box, _, _ = run_script(tag)
        prob, obj = box["prob"], box["obj"]
        ir = tape.build_ir(prob, obj)
        lb, ub = prob.bounds_arrays()
        arrays = tape.ir_to_arrays(ir)
        arrays.update(x0=np.array(prob.p, dtype=float), lb=lb, ub=ub,
                      ftol=np.float64(box["options"].get("ftol", 1e-6)),
                      maxiter=np.int64(box["options"].get("maxiter", 25)))
        path = os.path.join(GOLD, "example_%s_ir.npz" % tag)
        np.savez_compressed(path, **arrays)
        print("example", tag, "IR: nvars", ir.nvars, "tape words", [len(t.code) for t in ir.node_tapes],
              len(ir.scalar_tape.code), "tables", len(ir.tables), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
