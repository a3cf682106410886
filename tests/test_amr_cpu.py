"""Per-level chemistry over the synthetic TripleFlame hierarchy (BASELINE.json configs[4]): host logic of
pelelm_b200/amr.py -- costs, knapsack, box motion, average-down -- under gloo with world_size 2 on CPU, the integration
done by the oracle port standing in for the CUDA library (test infrastructure only; on the GPU the same driver calls
Reactor.advance_chemistry, tests/test_gpu_parity.py::test_amr_levels_vs_port)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

MECH = "LiDryer"
N0, BOX, NLEV, DT, NSTEPS = (4, 8, 4), 4, 3, 1e-5, 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _port_advance():
    from oracle import pyoracle
    P = pyoracle.Port(MECH).init(rtol=1e-8, atol=1e-10, max_order=2)
    P.set_typ_vals(np.full(P.n, -1.0))
    ns = P.n - 1

    def advance_box(bx, shape, dt, slot):
        S, F = bx["S"].numpy(), bx["F"].numpy()
        rY, rH, T, FC, I_R = pyoracle.advance_chemistry_soa(P, S[:ns].copy(), S[ns].copy(), S[ns + 1].copy(), F[:ns],
                                                            F[ns], bx["mask"].numpy(), dt)
        bx["S"][:ns] = torch.from_numpy(rY)
        bx["S"][ns] = torch.from_numpy(rH)
        bx["S"][ns + 1] = torch.from_numpy(T)
        bx["FC"].copy_(torch.from_numpy(FC))
        bx["I_R"].copy_(torch.from_numpy(I_R))
    return advance_box


def _run(rank, ws, knapsack):
    from pelelm_b200 import amr
    levels = amr.build_tripleflame(MECH, ws, rank, n0=N0, nlev=NLEV, box=BOX, forcing=True, dt=DT)
    adv = _port_advance()
    infos = []
    for _ in range(NSTEPS):
        infos.append(amr.advance_hierarchy(levels, DT, adv, use_knapsack=knapsack))
    out = {}
    for L in levels:
        for b, bx in L.data.items():
            out[(L.lev, b)] = {k: bx[k].numpy().copy() for k in ("S", "I_R", "I_R_old", "FC")}
    meta = [[(i["chem_owner"].tolist(), i["costs"].tolist(), i["efficiency"], i["moved"]) for i in step] for step in infos]
    return out, meta


def _worker(rank, ws, port, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    out, meta = _run(rank, ws, True)
    q.put((rank, out, meta))
    dist.barrier()
    dist.destroy_process_group()


def test_hierarchy_shape():
    from pelelm_b200 import amr
    lv = amr.tripleflame_boxes((32, 64, 32), 3, 32)
    assert [d for d, _ in lv] == [(32, 64, 32), (64, 128, 64), (128, 256, 128)]
    n = [len(b) for _, b in lv]
    assert n[0] == 2 and n[1] > n[0] and n[2] > n[1]
    # every fine grid is nested in the coarsened footprint of the next coarser level's grids, and grids do not overlap
    for l in range(1, 3):
        crs = {(lo[0] // 32, lo[1] // 32, lo[2] // 32) for lo, _ in lv[l - 1][1]}
        seen = set()
        for lo, hi in lv[l][1]:
            assert all((hi[d] - lo[d] + 1) == 32 and lo[d] % 32 == 0 for d in range(3))
            assert (lo[0] // 64, lo[1] // 64, lo[2] // 64) in crs
            assert lo not in seen
            seen.add(lo)
    # the refined region follows the mixing layer at x = 0.6 L
    xs = {lo[0] for lo, _ in lv[2][1]}
    assert any(abs((x + 16) / 128.0 - 0.6) < 0.15 for x in xs)


def test_two_ranks_match_one_rank_and_balance(port_lib):
    ws = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(ws)], key=lambda r: r[0])
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    one, meta1 = _run(0, 1, False)     # the same hierarchy in one process, default map
    got = {}
    got.update(res[0][1])
    got.update(res[1][1])
    assert sorted(got) == sorted(one), "every box is owned by exactly one rank"
    assert not (set(res[0][1]) & set(res[1][1]))
    for key in one:
        for name in ("S", "I_R", "I_R_old", "FC"):
            assert np.array_equal(got[key][name], one[key][name]), (key, name)   # the distribution changes nothing
    m0, m1 = res[0][2], res[1][2]
    assert m0 == m1, "ranks disagree on costs / chemistry map"
    # first pass: FC = 1 per cell -> costs = cells per box; later passes: measured FC, knapsack balances them
    from pelelm_b200 import amr
    lv = amr.tripleflame_boxes(N0, NLEV, BOX)
    for l in range(NLEV):
        assert m0[0][l][1] == [BOX ** 3] * len(lv[l][1])
        own, costs, eff, _ = m0[-1][l]
        assert max(costs) > BOX ** 3
        load = np.bincount(own, weights=costs, minlength=ws)
        if len(costs) > ws:   # <= nprocs boxes: the reference deals them round robin and reports efficiency 1
            assert eff is not None and abs(eff - load.mean() / load.max()) < 1e-12
        if len(costs) >= 4 * ws:
            assert eff >= 0.9
    assert any(step[l][3] > 0 for step in m0 for l in range(NLEV)), "the knapsack never moved a box: nothing was tested"


def test_average_down_of_covered_cells(port_lib):
    from pelelm_b200 import amr
    levels = amr.build_tripleflame(MECH, 1, 0, n0=N0, nlev=NLEV, box=BOX, forcing=True, dt=DT)
    adv = _port_advance()
    for _ in range(3):
        amr.advance_hierarchy(levels, DT, adv, use_knapsack=False)
    L, fine = levels[0], levels[1]
    ns = L.data[0]["I_R"].shape[0]
    # What level 0 must hold on covered cells after its 3rd advance: the mean over 2^3 fine cells of level 1's OLD I_R at
    # that moment.  Level 1 had not yet swapped its time levels for step 3 (advance_setup is per level, coarse level
    # first), so its old data was the result of its 1st advance: recompute that from a fresh one-step run.
    levels2 = amr.build_tripleflame(MECH, 1, 0, n0=N0, nlev=NLEV, box=BOX, forcing=True, dt=DT)
    amr.advance_hierarchy(levels2, DT, adv, use_knapsack=False)
    fine_prev = {b: bx["I_R"].numpy().copy() for b, bx in levels2[1].data.items()}
    assert any(np.abs(v).max() > 0 for v in fine_prev.values())
    nx, ny, nz = L.domain
    covered = np.zeros((nz, ny, nx), dtype=bool)
    want = np.zeros((ns, nz, ny, nx))
    for f, (lo, hi) in enumerate(fine.boxes):
        fx, fy, fz = fine.shape(f)
        v = fine_prev[f].reshape(ns, fz // 2, 2, fy // 2, 2, fx // 2, 2).mean(axis=(2, 4, 6))
        sl = (slice(lo[2] // 2, hi[2] // 2 + 1), slice(lo[1] // 2, hi[1] // 2 + 1), slice(lo[0] // 2, hi[0] // 2 + 1))
        want[(slice(None),) + sl] = v
        covered[sl] = True
    assert covered.any() and not covered.all()
    for b, (lo, hi) in enumerate(L.boxes):
        bx_, by_, bz_ = L.shape(b)
        got = L.data[b]["I_R"].numpy().reshape(ns, bz_, by_, bx_)
        sl = (slice(lo[2], hi[2] + 1), slice(lo[1], hi[1] + 1), slice(lo[0], hi[0] + 1))
        cov = covered[sl]
        assert np.array_equal(got[:, cov], want[(slice(None),) + sl][:, cov])
        # uncovered cells keep the level's own I_R = (new - old)/dt - F: non-zero wherever chemistry or forcing acts
        assert np.isfinite(got[:, ~cov]).all()
