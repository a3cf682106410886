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


def rel_errors(rY, rYr, atol=1e-10):
    """Measured relative errors of the species (what BASELINE.md section 3.4 asks to be stated): max |d rhoY_k| / rhoY_k over
    (a) every species with Y_k > 1e-12 and (b) the species above the integrator's own absolute tolerance (rhoY_k > atol),
    NaN cells excluded."""
    ok = ~(np.isnan(rYr).any(0) | np.isnan(rY).any(0))
    a, r = rY[:, ok], rYr[:, ok]
    if r.size == 0:
        return 0.0, 0.0
    Y = r / r.sum(0, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.abs(a - r) / np.abs(r)
    s1, s2 = Y > 1e-12, np.abs(r) > atol
    return (float(rel[s1].max()) if s1.any() else 0.0), (float(rel[s2].max()) if s2.any() else 0.0)


def check_state(got, ref, atol, tag="", rtol=1e-6, fc_equal_min=None, fc_rel=None):
    """got/ref = (rY, T, rE, FC). Parity bar of BASELINE.json north_star: final Y_k and T relative error <= 1e-6 at
    rtol = atol = 1e-10: |d rhoY_k| <= 1e-6 |rhoY_k| + atol (a species below the integrator's own absolute tolerance is
    compared against that tolerance: two correct integrations agree on it no better -- the reference's CPU build and its
    own restatement differ by up to 1e-2 relative on Y_k ~ 1e-9 where their step sequences differ in one step), rhoH
    bit-exact (it is in + dt*src in both), NaN cells (RK64 blow-up) identical, skipped cells identical."""
    rY, T, rE, FC = got
    rYr, Tr, rEr, FCr = ref
    nan_r = np.isnan(rYr).any(0) | np.isnan(Tr)
    nan_g = np.isnan(rY).any(0) | np.isnan(T)
    assert np.array_equal(nan_r, nan_g), tag + ": NaN cell pattern differs"
    ok = ~nan_r
    # the stated bar (floor = atol) holds at the tolerances it is stated for; where the integrator itself runs at a looser
    # rtol (comparison rtol = 100 x integrator rtol > 1e-6) two valid step sequences differ by a few integrator tolerances
    e = scaled_err(rY[:, ok], rYr[:, ok], rtol, atol if rtol <= 1e-6 else 10 * atol)
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
