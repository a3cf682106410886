#!/usr/bin/env python
"""A/B of the two kernel families on one GPU: right-hand sides per second and whole-cell advances
per second, warp-per-cell (path 1) against lane = cell (path 2), same states, results compared. Writes one JSON object."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pelelm_b200 import capi, synth  # noqa: E402

dev = torch.device("cuda", 0)
out = {}
nrhs = int(os.environ.get("AB_NRHS", 1 << 18))
ncell = int(os.environ.get("AB_NCELL", 1 << 16))
for mech in os.environ.get("AB_MECHS", "LiDryer,drm19").split(","):
    b = synth.synthetic_batch(mech, nrhs, seed=11, forcing=True)
    ns = b["rY"].shape[0]
    y = torch.from_numpy(np.ascontiguousarray(np.vstack([b["rY"], b["T"][None, :]]).T)).to(dev)
    rys = torch.from_numpy(np.ascontiguousarray(b["rYsrc"].T)).to(dev)
    rE, rEs = torch.from_numpy(b["rE"]).to(dev), torch.from_numpy(b["rEsrc"]).to(dev)
    res = {}
    for name, path in (("warp", 1), ("lane", 2)):
        R = capi.Reactor(mech, "ReactorCvode").init(rtol=1e-10, atol=1e-10, max_order=2).set_path(path)
        f = R.eval_rhs(y, rE, rEs, rys, t=1e-6)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(4):
            e0.record()
            R.eval_rhs(y, rE, rEs, rys, t=1e-6)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        J = R.eval_jac(y[:4096])
        torch.cuda.synchronize()
        bj = 1e30
        for _ in range(3):
            e0.record()
            R.eval_jac(y[:65536])
            e1.record()
            torch.cuda.synchronize()
            bj = min(bj, e0.elapsed_time(e1))
        res[name] = {"rhs_per_s": nrhs / (best * 1e-3), "rhs_ms": best, "jac_per_s": 65536 / (bj * 1e-3), "f": f.cpu().numpy(),
                     "J": J.cpu().numpy()}
        R.close()
    d = np.abs(res["lane"]["f"] - res["warp"]["f"]) / np.abs(res["warp"]["f"]).max(1, keepdims=True)
    dj = np.abs(res["lane"]["J"] - res["warp"]["J"]) / np.abs(res["warp"]["J"]).max(1, keepdims=True)
    out[mech] = {"n_rhs": nrhs, "rhs_per_s": {k: res[k]["rhs_per_s"] for k in res}, "jac_per_s": {k: res[k]["jac_per_s"] for k in res},
                 "rhs_speedup": res["lane"]["rhs_per_s"] / res["warp"]["rhs_per_s"],
                 "jac_speedup": res["lane"]["jac_per_s"] / res["warp"]["jac_per_s"],
                 "rhs_maxdiff_over_max": float(d.max()), "jac_maxdiff_over_max": float(dj.max())}
    # whole-cell advance
    for integ, tr in (("ReactorRK64", (900.0, 1150.0)), ("ReactorCvode", (900.0, 2600.0))):
        bb = synth.synthetic_batch(mech, ncell, seed=3, forcing=True, mask_every=9, t_range=tr)
        r2 = {}
        for name, path in (("warp", 1), ("lane", 2)):
            R = capi.Reactor(mech, integ).init(rtol=1e-10, atol=1e-10, max_order=2).set_path(path)
            best, st = 1e30, None
            for rep in range(2):
                a = [torch.from_numpy(bb[k].copy()).to(dev) for k in ("rY", "rYsrc", "T", "rE", "rEsrc")]
                FC = torch.zeros(ncell, dtype=torch.float64, device=dev)
                mk = torch.from_numpy(bb["mask"]).to(dev)
                st = R.react([0, 0, 0], [ncell - 1, 0, 0], a[0], a[1], a[2], a[3], a[4], FC, mk, 1e-5)
                best = min(best, st["kernel_ms"])
            r2[name] = {"cells_per_s": ncell / (best * 1e-3), "kernel_ms": best, "rY": a[0].cpu().numpy(), "T": a[2].cpu().numpy(),
                        "FC": FC.cpu().numpy(), "st": {k: st[k] for k in ("nst", "nfe", "nje", "nsetups", "nni", "netf", "ncfn", "nfailed")}}
            R.close()
        ok = ~np.isnan(r2["warp"]["rY"]).any(0)
        rel = np.abs(r2["lane"]["rY"] - r2["warp"]["rY"])[:, ok] / (1e-6 * np.abs(r2["warp"]["rY"][:, ok]) + 1e-10)
        out[mech][integ] = {"ncells": ncell, "cells_per_s": {k: r2[k]["cells_per_s"] for k in r2},
                            "speedup": r2["lane"]["cells_per_s"] / r2["warp"]["cells_per_s"],
                            "nan_same": bool(np.array_equal(ok, ~np.isnan(r2["lane"]["rY"]).any(0))),
                            "scaled_err_lane_vs_warp": float(rel.max()) if ok.any() else None,
                            "T_rel": float(np.max(np.abs(r2["lane"]["T"][ok] / r2["warp"]["T"][ok] - 1))) if ok.any() else None,
                            "fc_identical_frac": float(np.mean(r2["lane"]["FC"] == r2["warp"]["FC"])),
                            "stats": {k: r2[k]["st"] for k in r2}}
    print(mech, json.dumps(out[mech]), flush=True)
with open(os.path.join(ROOT, "gpurun_out", os.environ.get("AB_OUT", "r02a_lane_ab.json")), "w") as f:
    json.dump(out, f, indent=1)
