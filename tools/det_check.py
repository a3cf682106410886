"""Development: write the result of a small drm19 / LiDryer run to an npz (to compare native vs sanitizer vs occupancy-limited runs bit for bit)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pelelm_b200 import capi, synth
out = {}
for mech, integ, nc in (("drm19", "ReactorCvode", 64), ("LiDryer", "ReactorRK64", 64), ("LiDryer", "ReactorCvode", 64)):
    b = synth.synthetic_batch(mech, nc, seed=123, forcing=True, mask_every=9,
                              t_range=(900.0, 2600.0) if integ == "ReactorCvode" else (900.0, 1150.0))
    R = capi.Reactor(mech, integ).init(rtol=1e-10, atol=1e-10, max_order=2)
    a = [b["rY"].copy(), b["rYsrc"].copy(), b["T"].copy(), b["rE"].copy(), b["rEsrc"].copy(), np.zeros(nc)]
    R.react([0, 0, 0], [nc - 1, 0, 0], a[0], a[1], a[2], a[3], a[4], a[5], b["mask"], 1e-5)
    R.close()
    k = mech + "_" + integ
    out[k + "_rY"], out[k + "_T"], out[k + "_FC"] = a[0], a[2], a[5]
np.savez(sys.argv[1], **out)
