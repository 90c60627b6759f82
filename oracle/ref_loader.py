"""Load the *real* reference (`/root/reference/OpenGoddard/optimize.py`) under an
alias module name, unmodified, with the two environment shims it needs.

TEST INFRASTRUCTURE ONLY.  This only works where `/root/reference` exists (the
build container); it is used by `oracle/make_golden.py` to generate the golden
vectors under `tests/golden/` and by container-only tests that pin the numpy
restatement (`oracle/og_numpy.py`) against the reference itself.  Nothing that
runs on the GPU box may call it.

Shims:
  1. `matplotlib` is not installed -> chainable stub modules in `sys.modules`
     (reference imports it at OpenGoddard/optimize.py:35).
  2. `scipy.special.lpn` was removed in SciPy 1.15+ -> re-provide it from
     `scipy.special.legendre_p_all(n, x, diff_n=1)` (used at optimize.py:75,79).
"""
import importlib.util
import os
import sys
import types

def _find_reference():
    """$OPENGODDARD_REF, then /root/reference (the build container), then baseline/_ref (the
    unmodified reference package pip-installed there with --target; it is git-ignored but travels
    to the GPU box with the working tree, so bench.py's reference arm can time the real thing)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cands = [os.environ.get("OPENGODDARD_REF"), "/root/reference", os.path.join(here, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "OpenGoddard", "optimize.py")):
            return c
    return cands[0] or "/root/reference"


REFERENCE_ROOT = _find_reference()


class _Chain:
    """Chainable no-op object: every attribute / call / index returns itself."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return self

    def __call__(self, *a, **k):
        return self

    def __getitem__(self, k):
        return self

    def __iter__(self):
        return iter(())

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Chain()


def install_matplotlib_stub():
    """Put a chainable fake `matplotlib` / `matplotlib.pyplot` in sys.modules
    (only if the real one cannot be imported)."""
    try:
        import matplotlib.pyplot  # noqa: F401
        return False
    except Exception:
        pass
    top = _StubModule("matplotlib")
    top.__path__ = []
    pyplot = _StubModule("matplotlib.pyplot")
    top.pyplot = pyplot
    sys.modules["matplotlib"] = top
    sys.modules["matplotlib.pyplot"] = pyplot
    return True


def install_scipy_shims():
    from scipy import special, integrate
    if not hasattr(special, "lpn"):
        def lpn(n, x):
            out = special.legendre_p_all(n, x, diff_n=1)
            return out[0], out[1]
        special.lpn = lpn
    if not hasattr(integrate, "cumtrapz"):
        integrate.cumtrapz = integrate.cumulative_trapezoid


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "OpenGoddard", "optimize.py"))


_cached = None


def load_reference():
    """Return the reference's `optimize` module object (alias `og_ref_optimize`)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    install_matplotlib_stub()
    install_scipy_shims()
    path = os.path.join(REFERENCE_ROOT, "OpenGoddard", "optimize.py")
    spec = importlib.util.spec_from_file_location("og_ref_optimize", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cached = mod
    return mod


class Captured:
    """What the reference's `Problem.solve` handed to `scipy.optimize.minimize`
    (OpenGoddard/optimize.py:740-749): the real closures, unrestated."""

    def __init__(self, fun, x0, args, bounds, constraints, jac, options):
        self.fun, self.x0, self.args, self.bounds = fun, x0, args, bounds
        self.constraints, self.jac, self.options = constraints, jac, options

    def eq(self, x):
        c = self.constraints[0]
        return c["fun"](x, *c["args"])

    def ineq(self, x):
        c = self.constraints[1]
        return c["fun"](x, *c["args"])

    def cost(self, x):
        return self.fun(x, *self.args)


def capture_solve(mod, prob, obj):
    """Run `prob.solve(obj)` of the reference with `optimize.minimize` replaced by
    a recorder; returns the Captured closures.  `solve` exits after one pass
    because the recorder reports status 0 (optimize.py:753-754)."""
    import contextlib
    import io
    box = {}

    def recorder(fun, x0, args=(), bounds=None, constraints=(), jac=None,
                 method=None, options=None, **kw):
        box["cap"] = Captured(fun, x0.copy(), args, bounds, constraints, jac, options)
        return types.SimpleNamespace(message="captured", status=0, x=x0)

    real = mod.optimize.minimize
    # `mod.optimize` is scipy.optimize itself; patch a shim namespace instead
    shim = types.SimpleNamespace(minimize=recorder, root=mod.optimize.root)
    saved = mod.optimize
    mod.optimize = shim
    try:
        it = prob.iterator
        with contextlib.redirect_stdout(io.StringIO()):
            prob.solve(obj)
        prob.iterator = it
    finally:
        mod.optimize = saved
    assert real is saved.minimize
    return box["cap"]
