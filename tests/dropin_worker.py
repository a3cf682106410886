"""Run by tests/test_gpu_dropin.py in a subprocess (the reference keeps process-wide singletons: one mechanism per
process): the reference's own driver + factory create the B200 reactor BY NAME and the stock CPU reactor, both react
on the same box, results are written to an .npz."""
import sys

import numpy as np

sys.path.insert(0, sys.argv[1])
from oracle.pyoracle import Ref  # noqa: E402
from pelelm_b200 import synth  # noqa: E402

mech, integ, out = sys.argv[2], sys.argv[3], sys.argv[4]
nc = {"LiDryer": 96, "drm19": 96, "dodecane_lu": 8}[mech]
mild = integ == "ReactorRK64"
b = synth.synthetic_batch(mech, nc, seed=2718, forcing=True, mask_every=13, t_range=(900.0, 1150.0) if mild else (900.0, 2600.0))
R = Ref(mech, b200=True)
res = {}
for name in (integ, integ + "B200"):
    R.init(integrator=name, rtol=1e-10, atol=1e-10, max_order=2)
    R.set_typ_vals(np.full(R.n, -1.0))
    rY, T, rE, FC = R.react_soa(b["rY"], b["rYsrc"], b["T"], b["rE"], b["rEsrc"], b["mask"], 1e-5)
    res[name + "/rY"], res[name + "/T"], res[name + "/rE"], res[name + "/FC"] = rY, T, rE, FC
res["mask"] = b["mask"]
np.savez(out, **res)
