"""Shared parity helpers and the tolerance definitions used by every test."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def stacked_reference(g):
    """golden workload file -> c_ref (B, M) and J_ref (B, M, n) in [c_eq; c_ineq; cost] order."""
    c = np.concatenate((g["c_eq"], g["c_ineq"], g["cost"][:, None]), axis=1)
    J = np.concatenate((g["J_eq"], g["J_ineq"], g["g_cost"][:, None, :]), axis=1)
    return c, J


# ---- tolerance definitions (written here once, used by every parity test)
C_RTOL = 1e-6          # |c - c_ref| <= C_RTOL * max(|c_ref|, 1e-6 * terms_r): 1e-6 relative per entry;
#                        where a row is the cancellation of much larger terms (collocation defects:
#                        D.x ~ 1e2..1e3 minus the dynamics, leaving ~1e-6) the floor is 1e-12 of the
#                        first-order size of those terms, terms_r = sum_j |dc_r/dx_j| * |x_j|
J_RTOL = 1e-6          # |J - J_ref| <= J_RTOL * max_j |J_ref[r, j]|   (row-scaled) ...
#                        ... and the structural zero pattern must be identical
LGL_RTOL = 1e-8        # |a - b| <= LGL_RTOL * |b| (+ 1e-14 absolute for the exact-zero middle node)


def assert_c_close(c, c_ref, J_ref=None, x=None):
    """c, c_ref: (..., M); J_ref (..., M, n) and x (..., n) give the per-row term size."""
    c, c_ref = np.asarray(c), np.asarray(c_ref)
    assert c.shape == c_ref.shape
    if J_ref is not None:
        terms = (np.abs(J_ref) * np.abs(np.asarray(x))[..., None, :]).sum(axis=-1)
    else:
        terms = np.broadcast_to(np.abs(c_ref).max(axis=-1, keepdims=True), c_ref.shape)
    scale = np.maximum(np.abs(c_ref), 1e-6 * terms)
    bad = np.abs(c - c_ref) > C_RTOL * scale
    assert not bad.any(), "constraint vector mismatch: max rel %g" % (
        np.abs(c - c_ref) / np.maximum(scale, 1e-300)).max()


DUST_LOG = []          # (label, dust-level zero-pattern differences, entries compared): printed in pytest's terminal summary


J_DUST = 1e-7          # zero-pattern exceptions: a forward difference of two values that differ by an
#                        ulp or not at all (|entry| <= J_DUST * rowmax) may be 0 on one side only, because
#                        CUDA's and glibc's exp/sin/cos round differently in the last place


def assert_J_close(J, J_ref, dust=None, label=None):
    """J, J_ref: (..., M, n).  Row-scaled tolerance + identical zero pattern (up to FD dust:
    J_DUST * rowmax unless `dust` widens it, never beyond the value tolerance J_RTOL)."""
    dust = J_DUST if dust is None else min(float(dust), J_RTOL)
    J, J_ref = np.asarray(J), np.asarray(J_ref)
    assert J.shape == J_ref.shape
    rowmax = np.abs(J_ref).max(axis=-1, keepdims=True)
    err = np.abs(J - J_ref)
    bad = err > J_RTOL * rowmax
    assert not bad.any(), "Jacobian mismatch: max row-scaled err %g" % (
        err / np.maximum(rowmax, 1e-300)).max()
    mism = (J == 0) != (J_ref == 0)
    if label is None:
        import os
        label = os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0]
    DUST_LOG.append((label, int(mism.sum()), int(J.size)))
    if mism.any():
        big = np.maximum(np.abs(J), np.abs(J_ref)) > dust * rowmax
        assert not (mism & big).any(), "Jacobian zero pattern differs in %d entries (largest %g of rowmax)" % (
            (mism & big).sum(), (np.maximum(np.abs(J), np.abs(J_ref)) / np.maximum(rowmax, 1e-300))[mism].max())
        assert mism.sum() <= max(2, J.size // 100000), "too many dust-level zero-pattern differences: %d" % mism.sum()


def assert_lgl_close(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    assert (np.abs(a - b) <= LGL_RTOL * np.abs(b) + 1e-14).all(), "LGL mismatch %g" % np.abs(a - b).max()
