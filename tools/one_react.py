#!/usr/bin/env python
"""One react call on device arrays (for ncu captures): tools/one_react.py <mech> <integrator> <ncells> [path] [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pelelm_b200 import capi, synth  # noqa: E402

mech, integ, nc = sys.argv[1], sys.argv[2], int(sys.argv[3])
path = sys.argv[4] if len(sys.argv) > 4 else "auto"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
dev = torch.device("cuda", 0)
tr = (900.0, 1150.0) if integ == "ReactorRK64" else (900.0, 2600.0)
b = synth.synthetic_batch(mech, nc, seed=3, forcing=True, mask_every=9, t_range=tr)
R = capi.Reactor(mech, integ).init(rtol=1e-10, atol=1e-10, max_order=int(os.environ.get("MAX_ORDER", "2"))).set_path(path)
for _ in range(reps):
    a = [torch.from_numpy(b[k].copy()).to(dev) for k in ("rY", "rYsrc", "T", "rE", "rEsrc")]
    FC = torch.zeros(nc, dtype=torch.float64, device=dev)
    mk = torch.from_numpy(b["mask"]).to(dev)
    st = R.react([0, 0, 0], [nc - 1, 0, 0], a[0], a[1], a[2], a[3], a[4], FC, mk, 1e-5)
    print(mech, integ, path, nc, "cells: %.2f ms, %.0f cells/s" % (st["kernel_ms"], nc / st["kernel_ms"] * 1e3),
          {k: st[k] for k in ("nst", "nfe", "nsetups", "nje", "nfailed")}, flush=True)
R.close()
