"""The two Gauss-Jordan inverse routines of the CUDA integrator (bdf_warp.cuh: gjinv_warp, one pivot per sweep; gjinv2_warp,
two pivots per sweep, used for dodecane_lu) restated as a scalar C model (tests/gj_model.c) and compared on the CPU: same
inverse (up to the sign of exact zeros), same permutation, same singularity code -- the property that lets the two-pivot
routine replace the other without moving a single step-size decision.  The device routines themselves are covered by the
GPU parity tests; this pins the algorithm."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gj(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("gj") / "libgj.so")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", os.path.join(HERE, "gj_model.c"), "-o", so, "-lm"])
    L = C.CDLL(so)
    for f in (L.gj1, L.gj2):
        f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        f.restype = C.c_int
    return L


def run(f, A):
    n = A.shape[0]
    W = np.ascontiguousarray(A, dtype=np.float64).copy()
    perm = np.zeros(n, dtype=np.int32)
    rc = f(W.ctypes.data, n, n, perm.ctypes.data)
    return rc, W, perm


def same(gj, A, expect_rc=None):
    r1, W1, p1 = run(gj.gj1, A)
    r2, W2, p2 = run(gj.gj2, A)
    assert r1 == r2
    if expect_rc is not None:
        assert r1 == expect_rc
    if r1 == 0:
        assert np.array_equal(p1, p2)
        assert np.array_equal(W1, W2, equal_nan=True)   # -0.0 == +0.0: the sign of exact zeros may differ
    return r1, W1, p1


@pytest.mark.parametrize("n", [34, 54, 64])
def test_newton_like_matrices(gj, n):
    rng = np.random.default_rng(n)
    for trial in range(20):
        J = rng.standard_normal((n, n)) * np.exp(rng.uniform(-6, 6, (n, 1)))   # rows of very different scale, as a chemistry Jacobian
        A = np.eye(n) - 10.0 ** rng.uniform(-6, -1) * J
        rc, W, perm = same(gj, A, 0)
        assert np.abs(W @ A[perm] - np.eye(n)).max() < 1e-8


@pytest.mark.parametrize("n", [34, 54, 64])
def test_general_matrices_with_row_exchanges(gj, n):
    rng = np.random.default_rng(100 + n)
    swapped = 0
    for trial in range(20):
        A = rng.standard_normal((n, n))
        rc, W, perm = same(gj, A, 0)
        swapped += int(np.count_nonzero(perm != np.arange(n)))
        assert np.abs(W @ A[perm] - np.eye(n)).max() < 1e-6
    assert swapped > 10 * n


def test_ties_and_exact_zeros(gj):
    rng = np.random.default_rng(7)
    done = 0
    for trial in range(200):
        A = rng.integers(-2, 3, (54, 54)).astype(np.float64)   # many equal |a_ik|: the first row must win in both routines
        rc, W, perm = same(gj, A)
        done += rc == 0
    assert done > 100


@pytest.mark.parametrize("col", [0, 1, 20, 21, 52, 53])
def test_singular_column_reported_at_the_same_pivot(gj, col):
    rng = np.random.default_rng(col)
    A = rng.standard_normal((54, 54))
    A[:, col] = 0.0
    same(gj, A, col + 1)
    # a column that becomes (nearly or exactly) zero only during the elimination: two equal columns -- whatever rounding
    # leaves there, both routines must see the same thing
    B = rng.standard_normal((54, 54))
    if col > 0:
        B[:, col] = B[:, col - 1]
        same(gj, B)


def test_nan_never_wins_a_pivot_search(gj):
    rng = np.random.default_rng(3)
    A = rng.standard_normal((54, 54))
    A[40, 7] = np.nan
    r1, W1, p1 = run(gj.gj1, A)
    r2, W2, p2 = run(gj.gj2, A)
    assert r1 == r2 and np.array_equal(p1, p2) and np.array_equal(np.isnan(W1), np.isnan(W2))
