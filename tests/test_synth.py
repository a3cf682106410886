"""Synthetic-state generator: deterministic, shard-consistent, physically sane."""
import numpy as np
import pytest

from conftest import MECHS
from pelelm_b200 import synth


@pytest.mark.parametrize("mech", MECHS)
def test_shard_consistency(mech):
    full = synth.synthetic_batch(mech, 257, forcing=True, mask_every=97)
    a = synth.synthetic_batch(mech, 100, cell0=0, forcing=True, mask_every=97)
    b = synth.synthetic_batch(mech, 157, cell0=100, forcing=True, mask_every=97)
    for k in ("rY", "rYsrc"):
        assert np.array_equal(full[k], np.concatenate([a[k], b[k]], 1))
    for k in ("T", "rE", "rEsrc", "mask"):
        assert np.array_equal(full[k], np.concatenate([a[k], b[k]]))


@pytest.mark.parametrize("mech", MECHS)
def test_physical(mech):
    b = synth.synthetic_batch(mech, 512, forcing=True)
    m = synth.mech(mech)
    rho = b["rY"].sum(0)
    assert np.all(b["rY"] >= 0) and np.all(rho > 1e-5) and np.all(rho < 1e-2)
    assert 900 <= b["T_true"].min() and b["T_true"].max() <= 2600
    assert np.abs(b["rYsrc"].sum(0)).max() < 1e-9 * np.abs(b["rYsrc"]).max()
    # ideal gas at 1 atm
    p = rho * synth.RU * b["T_true"] * (b["rY"] / rho / m.mw[:, None]).sum(0)
    assert np.allclose(p, synth.P_ATM, rtol=1e-12)


def test_thermo_matches_oracle(port_lib):
    for mech in MECHS:
        b = synth.synthetic_batch(mech, 16)
        P = port_lib.Port(mech)
        rho = b["rY"].sum(0)
        Y = np.ascontiguousarray((b["rY"] / rho).T)
        r2, h2 = P.pty2rh(synth.P_ATM, b["T_true"], Y)
        assert np.allclose(r2, rho, rtol=1e-13)
        assert np.allclose(h2 * r2, b["rE"], rtol=1e-12)
