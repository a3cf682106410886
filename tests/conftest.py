"""Test configuration.

Markers: `gpu` = needs a B200 (run on the GPU box with `-m gpu`); everything else runs on CPU.
The checkers (oracle port, and the real reference through the committed golden fixtures) are test
infrastructure; the product path is pelelm_b200 + libplmchem.so, reached through the C ABI only.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MECHS = ["LiDryer", "drm19", "dodecane_lu"]
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def port_lib():
    from oracle import pyoracle
    pyoracle.build_port()
    return pyoracle


@pytest.fixture(scope="session")
def plm_lib():
    """libplmchem.so built in-tree (cross-compiles without a GPU)."""
    from pelelm_b200 import build, capi
    build.build_lib()
    return capi.load()


def golden(mech):
    return np.load(os.path.join(GOLDEN, mech + ".npz"))


def golden_plm(mech):
    """the PeleLM-side callers of the path (advance_chemistry body, reactionRateRhoY): tests/golden/<mech>_plm.npz"""
    return np.load(os.path.join(GOLDEN, mech + "_plm.npz"))


def golden_e(mech):
    """the internal-energy reactor (ode.reactor_type = 1): tests/golden/<mech>_e.npz"""
    return np.load(os.path.join(GOLDEN, mech + "_e.npz"))


def scaled_err(a, ref, rtol, atol):
    """max |a-ref| / (rtol*|ref| + atol): <= 1 means within the stated tolerance."""
    return float(np.max(np.abs(a - ref) / (rtol * np.abs(ref) + atol))) if ref.size else 0.0


def check_state(got, ref, atol, tag="", rtol=1e-6, fc_equal_min=None, fc_rel=None):
    """got/ref = (rY, T, rE, FC). Parity bar of BASELINE.json north_star: final Y_k and T relative error
    <= 1e-6 (species below the integrator's own absolute tolerance are compared against 10*atol), rhoH
    bit-exact (it is in + dt*src in both), NaN cells (RK64 blow-up) identical, skipped cells identical."""
    rY, T, rE, FC = got
    rYr, Tr, rEr, FCr = ref
    nan_r = np.isnan(rYr).any(0) | np.isnan(Tr)
    nan_g = np.isnan(rY).any(0) | np.isnan(T)
    assert np.array_equal(nan_r, nan_g), tag + ": NaN cell pattern differs"
    ok = ~nan_r
    e = scaled_err(rY[:, ok], rYr[:, ok], rtol, 10 * atol)
    assert e <= 1.0, "%s: rhoY scaled error %g" % (tag, e)
    et = scaled_err(T[ok], Tr[ok], rtol, 0.0)
    assert et <= 1.0, "%s: T relative error %g" % (tag, et * rtol)
    assert np.array_equal(rE[ok], rEr[ok]), tag + ": rhoH not bit-identical"
    assert np.array_equal(FC == 0, FCr == 0), tag + ": skipped-cell (FC=0) pattern differs"
    if fc_equal_min is not None and ok.any():
        frac = float(np.mean(FC[ok] == FCr[ok]))
        assert frac >= fc_equal_min, "%s: FC identical on only %.3f of cells" % (tag, frac)
    if fc_rel is not None and ok.any() and FCr[ok].sum() > 0:
        r = abs(FC[ok].sum() / FCr[ok].sum() - 1.0)
        assert r <= fc_rel, "%s: total FC differs by %.3f" % (tag, r)
    return e, et * rtol
