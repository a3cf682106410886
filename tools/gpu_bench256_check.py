#!/usr/bin/env python
"""Exercise bench.py's warm-up-free, replicated-box pass (the 256^3 entry of N = 8) at a small size on one GPU."""
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = [sys.argv[0]]
import bench  # noqa: E402

a = bench.parse()
ctx = bench.gpu_context()
sub = copy.copy(a)
sub.mech, sub.integrator, sub.cells, sub.box_cells = "dodecane_lu", "ReactorCvode", 32768, 8192
sub.steps, sub.warmup, sub.e2e_steps, sub.no_cpu_baseline, sub.allow_no_warmup, sub.replicate_boxes = 1, 0, 0, True, True, True
assert bench.fits_budget(ctx, 10.0, 800.0)
r = bench.run_batch(sub, ctx)
print(json.dumps({k: r[k] for k in ("value", "warmup", "steps", "ms_per_step", "gpu_launches")}), r["config"]["workload"], r["config"].get("warmup_note"))
