import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """on a machine without a CUDA device the gpu-marked tests are skipped (with the reason), not
    failed one by one; $OGB200_REQUIRE_GPU=1 turns that back into failures (a GPU box whose driver is broken)."""
    if os.environ.get("OGB200_REQUIRE_GPU") == "1" or not any("gpu" in it.keywords for it in items):
        return
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device on this machine (the hot path has no CPU fallback)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def api():
    import OpenGoddard.optimize as mod
    return mod


def pytest_terminal_summary(terminalreporter):
    """make the zero-pattern "dust" exceptions of the Jacobian comparisons visible -- entries that are 0
    on one side and a 1-ulp forward difference (<= 1e-7 of the row maximum) on the other (libm rounding)."""
    from tests import helpers
    if not helpers.DUST_LOG:
        return
    total = sum(n for _, n, _ in helpers.DUST_LOG)
    worst = sorted(helpers.DUST_LOG, key=lambda t: -t[1])[:5]
    terminalreporter.write_line("J zero-pattern dust: %d entries in %d comparisons (%d Jacobian entries); largest: %s" % (
        total, len(helpers.DUST_LOG), sum(sz for _, _, sz in helpers.DUST_LOG),
        ", ".join("%s=%d" % (lab.split("::")[-1], n) for lab, n, _ in worst if n) or "none"))
