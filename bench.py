#!/usr/bin/env python
"""Benchmark of the chemistry advance (BASELINE.json metric: chemistry cells*dt/s, % FP64 roofline).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--mech drm19|dodecane_lu|LiDryer]

Workload at N=1: BASELINE.json configs[2] -- a synthetic 2^22-cell drm19 batch (16 boxes of 64^3), ReactorCvode with
the dense analytic Jacobian, rtol = atol = 1e-10, max_order 2, dt = 1e-5 s, states from pelelm_b200/synth.py
(T 900-2600 K, phi 0.5-2, burnt fraction 0-1 with radicals).  N>1: the same per-GPU work (weak scaling), boxes assigned
by the reference's cost-weighted knapsack on the previous pass's FC; NCCL carries only the cost vector and the statistics.

The same JSON line carries, under "also", the other configurations BASELINE.json names, measured in the same run with the
same method on a bounded number of steps: dodecane_lu (configs[3]'s mechanism; 2^18 cells per GPU), LiDryer ReactorRK64
(configs[1], 2^20 cells per GPU) and the 3-level TripleFlame AMR hierarchy at the survey's size (configs[4]; strong scaling,
boxes re-dealt by the knapsack and moved between ranks over NCCL every pass).  --no-also skips them.  At N = 8 the dodecane_lu
entry is "dodecane_lu_256cubed" instead: configs[3] at its own size, 256^3 = 2^24 dodecane_lu cells over the 8 GPUs (one timed
pass of ~2 min without a warm-up pass of its own, every box of a rank holding the same 2^18 synthetic states so that the
synthesis is paid once; PLM_BENCH_256CUBED=1 adds it at another N).  Every nested workload runs only if it still fits the
invocation's wall-clock budget (PLM_BENCH_BUDGET_S = 790 s, the last one 800 s: the driver's scaling run allows 870 s per N
with --steps 20 --warmup 5); a workload that does not fit is reported as {"skipped": ...}.

A "step" = one react() pass over every box of the rank's shard, each step from the same pristine state (restored by a
device-to-device copy inside the timed region).  value = cells advanced by dt per second, whole job, inputs resident in
HBM; e2e = the same through the C ABI with HOST (pinned) arrays, host<->device copies inside the timed region.
roofline: FP64 SIMT (DFMA) -- algorithmic flops from the integrator counters (pelelm_b200/roofline.py) / step time,
against the DFMA peak measured live on this device.  cpu_baseline / --impl reference: the UNMODIFIED reference
(oracle/_ref/libref_<mech>.so, else the oracle port) on the host cores, one process per core, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DT = 1e-5
BOX = 64 ** 3
_T_IMPORT = time.time()
# Wall-clock budget of one invocation (seconds since the process started).  The driver gives a scaling run 870 s per N with
# --steps 20 --warmup 5 (SCALE_r01.json: per_n_timeout_s); a nested workload that does not fit is skipped and says so --
# a skipped entry costs one number, a run killed at the limit costs the whole line.
BUDGET_S = float(os.environ.get("PLM_BENCH_BUDGET_S", "790"))
BUDGET_LAST_S = float(os.environ.get("PLM_BENCH_BUDGET_LAST_S", "800"))   # the last entry: nothing comes after it
# cost of each nested workload (seconds, N = 1 incl. state synthesis, its e2e leg and CPU baseline, padded by ~20 %; the three
# of the default run took 138 s together in profiles/r02x_bench.json: ~70 + ~15 + ~53; the 256^3 pass is 115 s + 7 s of synthesis)
NESTED_COST_S = {"dodecane_lu": 80.0, "lidryer_rk64": 25.0, "tripleflame": 65.0, "dodecane_lu_256cubed": 140.0}


def elapsed_s():
    """seconds since this process was created (interpreter start-up and imports included)"""
    since_import = time.time() - _T_IMPORT
    try:
        import psutil
        e = time.time() - psutil.Process().create_time()
        if since_import <= e <= since_import + 600.0:   # a clock the container reports consistently
            return e
    except Exception:  # noqa: BLE001
        pass
    return since_import + 30.0


def fits_budget(ctx, cost_s, budget_s, now_s=None):
    """Does a nested workload of `cost_s` seconds still fit?  Rank 0 decides, every rank follows (the workloads contain
    collectives: all ranks must take the same branch)."""
    ok = ((elapsed_s() if now_s is None else now_s) + cost_s) <= budget_s
    if ctx is not None and ctx.get("world", 1) > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=ctx["dev"])
        dist.broadcast(t, src=0)
        ok = bool(int(t.item()))
    return ok


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mech", default="drm19")
    ap.add_argument("--integrator", default="ReactorCvode")
    ap.add_argument("--cells", type=int, default=0, help="cells per GPU (default 2^22 drm19, 2^20 LiDryer, 2^18 dodecane_lu)")
    ap.add_argument("--box-cells", type=int, default=BOX)
    ap.add_argument("--rtol", type=float, default=1e-10)
    ap.add_argument("--atol", type=float, default=1e-10)
    ap.add_argument("--max-order", type=int, default=2)
    ap.add_argument("--streams", type=int, default=4)
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target wall time of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="batch", choices=["batch", "tripleflame"],
                    help="batch: BASELINE.json configs[1-3] (default); tripleflame: configs[4], per-level chemistry of a synthetic "
                         "3-level AMR hierarchy with cost-weighted box distribution (pelelm_b200/amr.py)")
    ap.add_argument("--amr-n0", default="64,128,64", help="level-0 cells of the tripleflame hierarchy (SURVEY cfg5: 64,128,64)")
    ap.add_argument("--no-also", action="store_true", help="only the headline workload, not the nested ones")
    ap.add_argument("--also-steps", type=int, default=2, help="timed steps of each nested workload (bounded: the default run must end within minutes)")
    return ap.parse_args()


def default_cells(mech):
    return {"drm19": 1 << 22, "LiDryer": 1 << 20, "dodecane_lu": 1 << 18}[mech]


def workload_name(a, ncells_total, nboxes):
    return "synthetic %d-cell %s batch (%d boxes of %d cells), %s denseAJ_direct rtol=%g atol=%g max_order=%d dt=%g" % (
        ncells_total, a.mech, nboxes, a.box_cells, a.integrator, a.rtol, a.atol, a.max_order, DT)


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the reference itself (oracle/_ref) or the oracle port, one process per host core
# ---------------------------------------------------------------------------------------------------------------
_W = {}


def cpu_info():
    """model name, logical CPUs this process may use, and how many physical cores they are"""
    cpus = sorted(os.sched_getaffinity(0))
    model, phys, cur = "unknown", set(), {}
    try:
        for ln in open("/proc/cpuinfo"):
            if ":" not in ln:
                if cur and int(cur.get("processor", -1)) in cpus:
                    phys.add((cur.get("physical id", "0"), cur.get("core id", cur.get("processor"))))
                cur = {}
                continue
            k, v = [x.strip() for x in ln.split(":", 1)]
            cur[k] = v
            if k == "model name":
                model = v
        if cur and int(cur.get("processor", -1)) in cpus:
            phys.add((cur.get("physical id", "0"), cur.get("core id", cur.get("processor"))))
    except OSError:
        pass
    return {"cpu_model": model, "logical_cpus": len(cpus), "physical_cores": len(phys) or len(cpus)}


def _cpu_init(mech, integ, rtol, atol, max_order, kind, counter=None, cpus=None):
    from oracle import pyoracle
    if counter is not None and cpus:  # pin this worker to one logical CPU
        with counter.get_lock():
            i = counter.value
            counter.value += 1
        try:
            os.sched_setaffinity(0, {cpus[i % len(cpus)]})
        except OSError:
            pass
    if kind == "reference":
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)  # the reference prints its banner on stdout
        try:
            R = pyoracle.Ref(mech).init(integrator=integ, rtol=rtol, atol=atol, max_order=max_order)
        finally:
            os.dup2(saved, 1)
    else:
        R = pyoracle.Port(mech, build=False).init(integrator=integ, rtol=rtol, atol=atol, max_order=max_order)
    _W["R"] = R
    _W["mech"] = mech


def _cpu_chunk(arg):
    cell0, n = arg
    from pelelm_b200 import synth
    b = synth.synthetic_batch(_W["mech"], n, cell0=cell0)
    t0 = time.perf_counter()
    out = _W["R"].react_soa(b["rY"], b["rYsrc"], b["T"], b["rE"], b["rEsrc"], b["mask"], DT)
    return n, time.perf_counter() - t0, float(out[3].sum())


class CpuArm:
    def __init__(self, a):
        from oracle import pyoracle
        self.kind = "reference" if pyoracle.ref_available(a.mech) else "port"
        if self.kind == "port":
            pyoracle.build_port()
        cpus = sorted(os.sched_getaffinity(0))
        self.cores = len(cpus)
        self.info = cpu_info()
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.cores, initializer=_cpu_init,
                             initargs=(a.mech, a.integrator, a.rtol, a.atol, a.max_order, self.kind, ctx.Value("i", 0), cpus))
        self.a = a
        # calibrate: 32 cells on every core
        r = self.pool.map(_cpu_chunk, [(i * 32, 32) for i in range(self.cores)], chunksize=1)
        self.rate = self.cores * sum(x[0] for x in r) / sum(x[1] for x in r)  # cells/s over all cores

    def sample_cells(self, seconds):
        per_core = max(16, int(self.rate * seconds / self.cores))
        return per_core

    def run(self, per_core, cell0=0):
        chunks = []
        # 4 chunks per core so that a few expensive cells do not leave cores idle
        sub = max(4, per_core // 4)
        c = cell0
        for _ in range(self.cores * 4):
            chunks.append((c, sub))
            c += sub
        t0 = time.perf_counter()
        r = self.pool.map(_cpu_chunk, chunks, chunksize=1)
        wall = time.perf_counter() - t0
        n = sum(x[0] for x in r)
        return n, wall, sum(x[2] for x in r)

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm(a)
    cells_total = (a.cells or default_cells(a.mech)) * a.gpus
    nboxes = max(1, cells_total // a.box_cells)
    steps = max(1, a.steps)
    per_step_seconds = max(2.0, min(a.cpu_seconds, 150.0 / (steps + a.warmup)))
    per_core = arm.sample_cells(per_step_seconds)
    for _ in range(a.warmup):
        arm.run(max(16, per_core // 8))
    n_tot, wall_tot, fc_tot = 0, 0.0, 0.0
    for s in range(steps):
        n, wall, fc = arm.run(per_core, cell0=s * per_core * arm.cores)
        n_tot += n
        wall_tot += wall
        fc_tot += fc
    arm.close()
    value = n_tot / wall_tot
    sample = "%d consecutive cells of the same synthetic workload per step, %d processes" % (n_tot // steps, arm.cores)
    line = {
        "impl": "reference", "metric": "chemistry cells*dt/s", "value": value, "unit": "cells*dt/s", "n_gpus": a.gpus,
        "steps": steps, "warmup": a.warmup, "ms_per_step": 1e3 * wall_tot / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a, cells_total, nboxes), "cpu_sample": sample, "mean_fc": fc_tot / n_tot},
        "cpu_baseline": dict({"value": value, "unit": "cells*dt/s", "cores": arm.cores, "kind": arm.kind, "sample": sample,
                              "pinned": "one process per logical CPU (sched_setaffinity)"}, **arm.info),
        "e2e": {"value": value, "unit": "cells*dt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            t = [x.strip() for x in ln.split(",")]
            if len(t) < 7:
                continue
            try:
                sm.append(float(t[0]))
                mx.append(float(t[1]))
                pw.append(float(t[2]))
            except ValueError:
                continue
            for nm, v in zip(names, t[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "power_w_max": max(pw), "samples": len(sm)}
        return out


def gpu_context():
    import torch
    import torch.distributed as dist
    from pelelm_b200 import capi
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    capi.load()
    return {"rank": rank, "world": world, "local": local, "dev": dev}


def run_batch(a, ctx):
    """BASELINE.json configs[1-3]: returns the JSON line (rank 0) or None"""
    import torch
    import torch.distributed as dist
    from pelelm_b200 import capi, roofline, synth
    from pelelm_b200 import dist as pd

    rank, world, local, dev = ctx["rank"], ctx["world"], ctx["local"], ctx["dev"]
    N = world
    m = synth.mech(a.mech)
    ns = m.ns
    cells_per_gpu = a.cells or default_cells(a.mech)
    box_cells = min(a.box_cells, cells_per_gpu)
    nboxes = N * (cells_per_gpu // box_cells)
    cells_total = nboxes * box_cells
    owner = (np.arange(nboxes) % N).astype(np.int32)  # round-robin start, as getFuncCountDM (PeleLM.cpp:5537-5538)

    R = capi.Reactor(a.mech, a.integrator).init(rtol=a.rtol, atol=a.atol, max_order=a.max_order)
    peak_tf = capi.measure_fp64_peak()

    # ---- pristine state of this rank's boxes: S0 [ns+2][B] = rhoY, rhoH, T (the STemp aliasing of PeleLM.cpp:5728-5730),
    #      F [ns+1][B] = forcing (FTemp), mask [B]; pinned host copies feed the e2e leg
    rep = getattr(a, "replicate_boxes", False)   # every box of the rank = its first box (same cell-cost distribution; the
    first_local = int(np.nonzero(owner == rank)[0][0]) if rep and np.any(owner == rank) else 0   # synthesis is paid once)
    synth_cache = {}

    def make_box(b):
        key = first_local if rep else b
        if key not in synth_cache:
            synth_cache.clear()
            synth_cache[key] = synth.synthetic_batch(a.mech, box_cells, cell0=key * box_cells)
        st = synth_cache[key]
        S = torch.empty((ns + 2, box_cells), dtype=torch.float64, pin_memory=True)
        S[:ns] = torch.from_numpy(st["rY"])
        S[ns] = torch.from_numpy(st["rE"])
        S[ns + 1] = torch.from_numpy(st["T"])
        F = torch.zeros((ns + 1, box_cells), dtype=torch.float64, pin_memory=True)
        mask = torch.from_numpy(st["mask"]).pin_memory()
        return {"id": b, "hS": S, "hF": F, "hmask": mask}

    boxes = {}
    for b in np.nonzero(owner == rank)[0]:
        boxes[int(b)] = make_box(int(b))

    def to_device(bx):
        bx["S0"] = bx["hS"].to(dev, non_blocking=True)
        bx["F"] = bx["hF"].to(dev, non_blocking=True)
        bx["mask"] = bx["hmask"].to(dev, non_blocking=True)
        bx["S"] = torch.empty_like(bx["S0"])
        bx["FC"] = torch.zeros(box_cells, dtype=torch.float64, device=dev)

    for bx in boxes.values():
        to_device(bx)
    torch.cuda.synchronize()
    FCall = FCoff = None

    def bind_fc():
        """FC of all local boxes as views of one contiguous array, so that the per-box cost reduction is one launch."""
        nonlocal FCall, FCoff
        ids = sorted(boxes)
        FCall = torch.zeros(len(ids) * box_cells, dtype=torch.float64, device=dev)
        for i, b in enumerate(ids):
            boxes[b]["FC"] = FCall[i * box_cells:(i + 1) * box_cells]
        FCoff = torch.arange(0, (len(ids) + 1) * box_cells, box_cells, dtype=torch.int64, device=dev)

    bind_fc()

    streams = [torch.cuda.Stream(device=dev) for _ in range(max(1, a.streams))]
    main = torch.cuda.current_stream()
    launches = [0]

    def react_box(bx, s, host=False):
        lo, hi = [0, 0, 0], [box_cells - 1, 0, 0]
        if host:
            S, F, FC, mk = bx["hSw"].numpy(), bx["hF"].numpy(), bx["hFC"].numpy(), bx["hmask"].numpy()
        else:
            S, F, FC, mk = bx["S"], bx["F"], bx["FC"], bx["mask"]
        R.react(lo, hi, S[:ns], F[:ns], S[ns + 1], S[ns], F[ns], FC, mk, DT, stream=s.cuda_stream, sync=False)
        launches[0] += 1

    def step():
        """One advance over the local shard + the cost / statistics reductions of advance_chemistry."""
        ids = sorted(boxes)
        if not R.path_is_per_box():   # every local box in ONE launch (plm_react_boxes): the queue's tail is paid once per step
            lst = []
            for b in ids:
                bx = boxes[b]
                bx["S"].copy_(bx["S0"], non_blocking=True)
                S, F = bx["S"], bx["F"]
                lst.append({"lo": [0, 0, 0], "hi": [box_cells - 1, 0, 0], "rY": S[:ns], "rYsrc": F[:ns], "T": S[ns + 1], "rEner": S[ns],
                            "rEner_src": F[ns], "FC": bx["FC"], "mask": bx["mask"]})
            R.react_boxes(lst, DT, stream=main.cuda_stream, sync=False)
            launches[0] += 1
        else:
            for i, b in enumerate(ids):
                bx = boxes[b]
                s = streams[i % len(streams)]
                with torch.cuda.stream(s):
                    bx["S"].copy_(bx["S0"], non_blocking=True)
                    react_box(bx, s)
        st = R.stats()  # waits for every launch
        # per-box cost = sum of FC (PeleLM::getFuncCountDM): one launch of the library's reduction kernel over all local boxes
        costs = capi.box_costs(FCall, FCoff).cpu().numpy()
        launches[0] += 1
        full = pd.gather_box_costs(costs, owner)
        new_owner, eff = pd.func_count_dm(full, N, default_owner=owner)
        return pd.reduce_stats(st), full, new_owner, eff

    def barrier():
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up; after the first pass the boxes are re-dealt by the knapsack on the measured costs
    eff = 1.0
    nwarm = a.warmup if getattr(a, "allow_no_warmup", False) else max(a.warmup, 1)
    for w in range(nwarm):
        gst, full, new_owner, eff = step()
        if w == 0 and N > 1 and not np.array_equal(new_owner, owner):
            moved = pd.move_boxes(boxes, ("hS", "hF", "hmask"), owner, new_owner, device=dev)
            for b, bx in moved.items():
                if owner[b] != rank:  # received: back to pinned host memory, where the e2e leg reads it
                    for key in ("hS", "hF", "hmask"):
                        bx[key] = bx[key].cpu().pin_memory()
            boxes = moved
            owner = new_owner.copy()
            for bx in boxes.values():
                to_device(bx)
            bind_fc()
            torch.cuda.synchronize()

    # ---- timed region
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches[0] = 0
    e0.record(main)
    for s in streams:
        s.wait_event(e0)
    tot = None
    for _ in range(a.steps):
        gst, full, new_owner, eff = step()
        tot = gst if tot is None else {k: (tot[k] + gst[k] if k not in ("max_fc", "min_fc") else gst[k]) for k in gst}
    for s in streams:
        main.wait_stream(s)
    e1.record(main)
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if clocks else None
    if N > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    n_launch = launches[0] * N  # every rank: one integration kernel per local box per step + one box-cost reduction per step
    value = cells_total * a.steps / (ms * 1e-3)
    flops = roofline.algorithmic_flops(a.mech, a.integrator, tot)
    achieved = flops / (ms * 1e-3) * 1e-12 / N  # per GPU
    per_cell = {k: tot[k] / max(tot["ncells"], 1) for k in ("nst", "nfe", "nje", "nsetups", "nni", "netf", "ncfn")}

    # ---- e2e: same step through the C ABI with HOST arrays (pinned), copies inside the timed region
    e2e = None
    if a.e2e_steps > 0:
        for bx in boxes.values():
            bx["hSw"] = torch.empty_like(bx["hS"]).pin_memory()
            bx["hFC"] = torch.zeros(box_cells, dtype=torch.float64).pin_memory()
        es = streams[0]
        ms_e = 0.0
        for _ in range(a.e2e_steps):
            for bx in boxes.values():
                bx["hSw"].copy_(bx["hS"])
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(es)
            for b in sorted(boxes):
                react_box(boxes[b], es, host=True)   # H2D, kernel, D2H, sync -- inside the library
            R.stats()
            a1.record(es)
            barrier()
            ms_e += a0.elapsed_time(a1)
        if N > 1:
            t = torch.tensor([ms_e], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e = float(t.item())
        nloc = cells_per_gpu
        h2d = N * nloc * ((2 * ns + 4) * 8 + 4)   # rhoY, rhoH, T, forcing (N+1), FC staging, mask
        d2h = N * nloc * (ns + 3) * 8             # rhoY, rhoH, T, FC
        e2e = {"value": cells_total * a.e2e_steps / (ms_e * 1e-3), "unit": "cells*dt/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": a.e2e_steps}

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and N == 1 and not a.no_cpu_baseline:
        arm = CpuArm(a)
        per_core = arm.sample_cells(a.cpu_seconds)
        n, wall, _ = arm.run(per_core)
        arm.close()
        cpu = dict({"value": n / wall, "unit": "cells*dt/s", "cores": arm.cores, "kind": arm.kind,
                    "sample": "first %d cells of the same workload, %d processes, %.1f s" % (n, arm.cores, wall),
                    "pinned": "one process per logical CPU (sched_setaffinity)"}, **arm.info)

    line = None
    if rank == 0:
        # DRAM bytes per launch: NOT measured by this run (ncu cannot wrap the timed region); it is the per-cell figure of the
        # committed ncu capture of the same kernel (profiles/ncu_traffic.json names it) x the cells of one launch
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):
            try:
                ent = json.load(open(tp)).get(a.mech if a.integrator == "ReactorCvode" else a.mech + "_rk64", {})
                dram_per_cell = ent.get("dram_bytes_per_cell")
                # one launch = every local box of the rank (plm_react_boxes), or one box (PLM_BOXES_PER_LAUNCH=1)
                launch_cells = box_cells if R.path_is_per_box() else cells_per_gpu
                traffic = dram_per_cell * launch_cells if dram_per_cell else None
                traffic_src = ent.get("source")
            except Exception:
                traffic = None
        line = {
            "metric": "chemistry cells*dt/s", "value": value, "unit": "cells*dt/s", "n_gpus": N, "steps": a.steps,
            "warmup": nwarm, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a, cells_total, nboxes), "l2": "inputs larger than L2 (%.2f GB state per GPU)"
                       % (cells_per_gpu * roofline.state_bytes_per_cell(a.mech) / 1e9),
                       "cells_per_gpu": cells_per_gpu, "parallelism": "boxes over %d GPU(s), knapsack on FC" % N,
                       "knapsack_efficiency": eff, "per_cell": per_cell, "nfailed": tot["nfailed"],
                       "state_io_GBs": value / N * roofline.state_bytes_per_cell(a.mech) / 1e9},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf if peak_tf else None, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": "DFMA probe measured live on this device (MEASURED_PEAKS.json has no FP64 entry)",
                         "flops_per_cell": flops / max(tot["ncells"], 1)},
            "clocks": clk, "e2e": e2e, "gpu_launches": n_launch,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if rep:
            line["config"]["workload"] += "; every box of a rank holds the same %d synthetic states" % box_cells
        if nwarm == 0:
            line["config"]["warmup_note"] = ("no warm-up pass: one pass of this workload takes minutes, first-launch costs (module load, "
                                             "scratch allocation) are well under a second of it, and its equal-cost boxes give the knapsack nothing to move")
    R.close()
    del boxes
    torch.cuda.empty_cache()
    return line


def run_amr(a, ctx):
    """BASELINE.json configs[4]: one step = one PeleLM::advance_chemistry of every level of the synthetic TripleFlame
    hierarchy (dt_L = dt / 2^L), boxes dealt over the GPUs by the knapsack on the previous pass's FC (state map = round
    robin), box payloads moved to the chemistry owner and back, covered I_R averaged down.  Strong scaling: the hierarchy
    is the same for every N.  Returns the JSON line (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from pelelm_b200 import amr, capi, roofline
    from pelelm_b200 import dist as pd

    rank, world, local, dev = ctx["rank"], ctx["world"], ctx["local"], ctx["dev"]
    N = world
    n0 = tuple(int(x) for x in a.amr_n0.split(","))
    R = capi.Reactor(a.mech, a.integrator).init(rtol=a.rtol, atol=a.atol, max_order=a.max_order)
    peak_tf = capi.measure_fp64_peak()
    levels = amr.build_tripleflame(a.mech, N, rank, n0=n0, nlev=3, box=32, device=dev, dt=DT)
    for L in levels:
        for bx in L.data.values():
            bx["S0"] = bx["S"].clone()
    streams = [torch.cuda.Stream(device=dev) for _ in range(max(1, a.streams))]
    adv, wait_k = amr.gpu_advance(R, streams)
    tot = {}

    def wait():
        st = wait_k()
        for k, v in st.items():
            tot[k] = max(tot.get(k, 0), v) if k in ("max_fc", "min_fc", "kernel_ms") else tot.get(k, 0) + v

    def step():
        infos = []
        for L in levels:
            for bx in L.data.values():
                bx["S"].copy_(bx["S0"], non_blocking=True)   # every step from the same pristine state; FC is kept: the costs
            fine = levels[L.lev + 1] if L.lev + 1 < len(levels) else None
            infos.append(amr.advance_chemistry_level(L, DT / 2 ** L.lev, adv, fine=fine, cost_fn=capi.box_costs, wait=wait,
                                                     device=dev))
        return infos

    def barrier():
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 1)):
        infos = step()
    barrier()
    tot.clear()
    adv.launches[0] = 0
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        infos = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if clocks else None
    gst = pd.reduce_stats(tot)
    launches = adv.launches[0] + a.steps * sum(1 for L in levels if L.data)   # advances + one box-cost reduction per level
    if N > 1:
        t = torch.tensor([ms, float(launches)], dtype=torch.float64, device=dev)
        dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(t[1:], op=dist.ReduceOp.SUM)
        ms, launches = float(t[0].item()), int(t[1].item())
    line = None
    if rank == 0:
        ncell = [L.ncells() for L in levels]
        cells_total = sum(ncell)
        value = cells_total * a.steps / (ms * 1e-3)
        flops = roofline.algorithmic_flops(a.mech, a.integrator, gst)
        achieved = flops / (ms * 1e-3) * 1e-12 / N
        per_cell = {k: gst[k] / max(gst["ncells"], 1) for k in ("nst", "nfe", "nje", "nsetups", "nni", "netf", "ncfn")}
        line = {
            "metric": "chemistry cells*dt/s", "value": value, "unit": "cells*dt/s", "n_gpus": N, "steps": a.steps,
            "warmup": max(a.warmup, 1), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic 3-level TripleFlame AMR, %s %s denseAJ_direct rtol=%g atol=%g max_order=%d, level-0 %s cells, "
                                   "boxes of 32^3 per level %s, cells per level %s, dt_L = %g/2^L"
                                   % (a.mech, a.integrator, a.rtol, a.atol, a.max_order, "x".join(map(str, n0)),
                                      [len(L.boxes) for L in levels], ncell, DT),
                       "l2": "inputs larger than L2 (%.2f GB state)" % (cells_total * roofline.state_bytes_per_cell(a.mech) / 1e9),
                       "parallelism": "boxes of every level over %d GPU(s), knapsack on the previous pass's FC" % N,
                       "knapsack_efficiency": [i["efficiency"] for i in infos], "boxes_moved": [i["moved"] for i in infos],
                       "per_cell": per_cell, "nfailed": gst["nfailed"]},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf if peak_tf else None, "traffic": None,
                         "peak_source": "DFMA probe measured live on this device (MEASURED_PEAKS.json has no FP64 entry)",
                         "flops_per_cell": flops / max(gst["ncells"], 1)},
            "clocks": clk, "e2e": None, "gpu_launches": launches,
        }
    R.close()
    del levels
    torch.cuda.empty_cache()
    return line


NESTED_KEYS = ("value", "unit", "steps", "warmup", "ms_per_step", "scaling", "config", "roofline", "e2e", "cpu_baseline", "gpu_launches")


def run_gpu_arm(a):
    import copy
    ctx = gpu_context()
    if a.workload == "tripleflame":
        line = run_amr(a, ctx)
    else:
        line = run_batch(a, ctx)
        headline = (a.mech == "drm19" and a.integrator == "ReactorCvode" and not a.cells)
        if headline and not a.no_also:
            also = {}
            ks = max(1, min(a.steps, a.also_steps))
            want256 = ctx["world"] == 8 or os.environ.get("PLM_BENCH_256CUBED") == "1"
            for name, kw in (("dodecane_lu", dict(mech="dodecane_lu", integrator="ReactorCvode")),
                             ("lidryer_rk64", dict(mech="LiDryer", integrator="ReactorRK64")),
                             ("tripleflame", dict(workload="tripleflame"))):
                if name == "dodecane_lu" and ctx["world"] == 8:
                    continue   # N = 8: the same kernel runs the 256^3 domain below instead (no time for both in the driver's 870 s)
                sub = copy.copy(a)
                for k, v in kw.items():
                    setattr(sub, k, v)
                sub.steps, sub.warmup, sub.e2e_steps = ks, 1, (1 if name != "tripleflame" else 0)
                sub.cpu_seconds = min(a.cpu_seconds, 8.0)
                if not fits_budget(ctx, NESTED_COST_S[name], BUDGET_S):
                    also[name] = {"skipped": "time budget: %.0f s elapsed, %.0f s needed, %.0f s allowed (PLM_BENCH_BUDGET_S)"
                                  % (elapsed_s(), NESTED_COST_S[name], BUDGET_S)}
                    continue
                try:   # a nested workload that fails (on every rank alike) must not cost the headline its line
                    r = run_amr(sub, ctx) if name == "tripleflame" else run_batch(sub, ctx)
                except Exception as e:  # noqa: BLE001
                    r = None
                    also[name] = {"error": "%s: %s" % (type(e).__name__, e)}
                if r is not None:
                    also[name] = {k: r[k] for k in NESTED_KEYS if k in r}
            # BASELINE.json configs[3] at its own size: 256^3 dodecane_lu cells over the GPUs of the box.  One pass over 2^24
            # cells costs ~2 min on 8 B200 and proportionally more on fewer, so the full domain runs where the config is
            # named with the most GPUs (N = 8; PLM_BENCH_256CUBED=1 forces it at any N) and only if it still fits the time
            # budget; the 2^18-cells-per-GPU entry is the same kernel on the same states at the other N.  ONE timed pass without
            # a warm-up pass (first-launch costs are well under a second of two minutes; equal-cost boxes: the knapsack of a
            # warm-up pass would move nothing).  A failure here must not cost the line.
            if want256 and not fits_budget(ctx, NESTED_COST_S["dodecane_lu_256cubed"], BUDGET_LAST_S):
                also["dodecane_lu_256cubed"] = {"skipped": "time budget: %.0f s elapsed, %.0f s needed, %.0f s allowed "
                                                "(PLM_BENCH_BUDGET_LAST_S)" % (elapsed_s(), NESTED_COST_S["dodecane_lu_256cubed"], BUDGET_LAST_S)}
                want256 = False
            if want256:
                sub = copy.copy(a)
                sub.mech, sub.integrator = "dodecane_lu", "ReactorCvode"
                sub.cells = int(os.environ.get("PLM_BENCH_256CUBED_CELLS", 256 ** 3)) // ctx["world"]
                sub.steps, sub.warmup, sub.e2e_steps, sub.no_cpu_baseline, sub.allow_no_warmup = 1, 0, 0, True, True
                sub.replicate_boxes = True   # 7 s of state synthesis per rank instead of 58 s
                try:
                    r = run_batch(sub, ctx)
                    if r is not None:
                        also["dodecane_lu_256cubed"] = {k: r[k] for k in NESTED_KEYS if k in r}
                        also["dodecane_lu_256cubed"]["scaling"] = "strong"
                except Exception as e:  # noqa: BLE001
                    if line is not None:
                        also["dodecane_lu_256cubed"] = {"error": "%s: %s" % (type(e).__name__, e)}
            if line is not None:
                line["also"] = also
    if line is not None:
        print(json.dumps(line), flush=True)
    if ctx["world"] > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_gpu_arm(a)


if __name__ == "__main__":
    main()
