#!/usr/bin/env python
"""Structural fill of the chemistry Jacobian and of its LU factors (is a sparse-pattern
factorisation worth building?).

    python tools/jac_fill.py [mech ...]

The pattern is the union of the non-zeros of the analytic Jacobian (the unmodified reference's cJac through
oracle/_ref, else the oracle port) over synthetic states of the bench's range, i.e. what the reference's
SPARSITY_INFO_SYST counts.  For the Newton matrix I - gamma J it prints the fill, the fill of L + U after a symbolic
factorisation in the natural order (no pivoting: a lower bound on what partial pivoting leaves), the multiply-adds of a
factorisation that skips structural zeros against the dense 1/3 n^3, and the fill of the inverse (what the CUDA kernels
form: a structurally full matrix as soon as the graph is strongly connected).  Test infrastructure (uses oracle/)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402
from pelelm_b200 import synth  # noqa: E402


def pattern(mech, nstates=64):
    src = pyoracle.Ref(mech) if pyoracle.ref_available(mech) else pyoracle.Port(mech)
    src.init()
    b = synth.synthetic_batch(mech, nstates, seed=7, forcing=False, mask_every=0, t_range=(900.0, 2600.0))
    n = b["rY"].shape[0] + 1
    pat = np.zeros((n, n), dtype=bool)
    y = np.ascontiguousarray(np.concatenate([b["rY"], b["T"][None, :]], axis=0).T)  # [cell][n]
    J = np.asarray(src.jac(y)).reshape(nstates, n, n)  # column-major per cell, like the SUNDenseMatrix
    pat = (J != 0.0).any(axis=0).T
    return pat, type(src).__name__


def symbolic_lu(pat):
    a = pat.copy()
    n = a.shape[0]
    np.fill_diagonal(a, True)
    fma = 0
    for k in range(n):
        rows = np.nonzero(a[k + 1:, k])[0] + k + 1
        cols = np.nonzero(a[k, k + 1:])[0] + k + 1
        fma += len(rows) * len(cols)
        if len(rows) and len(cols):
            a[np.ix_(rows, cols)] = True
    return a, fma


def main():
    for mech in sys.argv[1:] or ["LiDryer", "drm19", "dodecane_lu"]:
        pat, who = pattern(mech)
        n = pat.shape[0]
        A = pat.copy()
        np.fill_diagonal(A, True)
        lu, fma = symbolic_lu(pat)
        # structural inverse: reachability (transitive closure) of the graph of A
        reach = A.copy()
        for k in range(n):
            reach |= np.outer(reach[:, k], reach[k, :])
        dense_fma = sum((n - 1 - k) ** 2 for k in range(n))
        # transposed (column-major like the reference's dense matrix: same counts)
        print("%-12s n = %2d (%s): J fill %5.1f %%, I - gamma J %5.1f %%, L + U after symbolic LU %5.1f %%, "
              "multiply-adds %d of %d dense (%.0f %%), inverse fill %5.1f %%"
              % (mech, n, who, 100.0 * pat.mean(), 100.0 * A.mean(), 100.0 * lu.mean(), fma, dense_fma,
                 100.0 * fma / dense_fma, 100.0 * reach.mean()))


if __name__ == "__main__":
    main()
