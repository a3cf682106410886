"""Drop-in check on the GPU box: the UNMODIFIED reference driver (oracle/_ref/libref_<mech>_b200.so = reference
objects + pelelm_b200/adaptor/ReactorB200.cpp) selects `chem_integrator = Reactor{Cvode,RK64}B200` through its own
factory (Factory.H:73-80) and calls the virtual react(box, ...) exactly as PeleLM does (PeleLM.cpp:5752-5758); the
result is compared with the stock CPU reactor run in the same process on the same FAB."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, check_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mech,integ", [("drm19", "ReactorCvode"), ("LiDryer", "ReactorCvode"),
                                        ("LiDryer", "ReactorRK64"), ("dodecane_lu", "ReactorCvode")])
def test_reference_factory_creates_b200_reactor(tmp_path, plm_lib, mech, integ):
    from oracle import pyoracle
    if not pyoracle.ref_available(mech, b200=True):
        pytest.skip("oracle/_ref/libref_%s_b200.so not built (needs /root/reference at build time)" % mech)
    out = str(tmp_path / "r.npz")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_worker.py"), ROOT, mech, integ, out],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    g = np.load(out)
    a, b = integ, integ + "B200"
    small = mech != "dodecane_lu"
    check_state((g[b + "/rY"], g[b + "/T"], g[b + "/rE"], g[b + "/FC"]),
                (g[a + "/rY"], g[a + "/T"], g[a + "/rE"], g[a + "/FC"]), 1e-10, tag="%s/%s" % (mech, integ),
                fc_equal_min=0.9 if small else None, fc_rel=0.05 if small else 0.15)
