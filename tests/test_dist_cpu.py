"""Host-side multi-rank logic on CPU: world_size 2 over gloo (127.0.0.1)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from pelelm_b200 import dist as pd
    nb = 13
    rng = np.random.default_rng(0)
    costs = rng.integers(1, 100000, nb).astype(np.int64)       # ground truth known to both ranks
    owner = (np.arange(nb) % ws).astype(np.int32)              # round-robin start (getFuncCountDM :5537-5538)
    mine = np.nonzero(owner == rank)[0]
    full = pd.gather_box_costs(costs[mine], owner)
    assert np.array_equal(full, costs)
    new_owner, eff = pd.func_count_dm(full, ws)
    # zero FC keeps the default mapping
    keep, _ = pd.func_count_dm(np.zeros(nb, dtype=np.int64), ws, default_owner=owner)
    assert np.array_equal(keep, owner)
    payload = {int(b): torch.full((3, 4), float(b)) for b in mine}
    got = pd.redistribute(payload, owner, new_owner)
    assert sorted(got) == sorted(int(b) for b in np.nonzero(new_owner == rank)[0])
    for b, t in got.items():
        assert torch.all(t == float(b))
    # whole boxes with several arrays each (what bench.py / a PeleLM driver moves after the first knapsack pass)
    boxes = {int(b): {"id": int(b), "S": torch.full((2, 5), float(b)), "F": torch.full((3,), -float(b)),
                      "m": torch.full((5,), int(b), dtype=torch.int32)} for b in mine}
    moved = pd.move_boxes(boxes, ("S", "F", "m"), owner, new_owner)
    assert sorted(moved) == sorted(int(b) for b in np.nonzero(new_owner == rank)[0])
    for b, bx in moved.items():
        assert torch.all(bx["S"] == float(b)) and torch.all(bx["F"] == -float(b)) and torch.all(bx["m"] == int(b))
        assert bx["m"].dtype == torch.int32 and tuple(bx["S"].shape) == (2, 5)
    # LeastUsedCPUs: one all-gather of the kilobytes each rank holds; the heaviest bucket goes to the least-used rank
    order = pd.least_used_ranks(5000 if rank == 0 else 100)
    assert order.tolist() == [1, 0]
    own2, eff2 = pd.func_count_dm(full, ws, rank_order=order)
    assert np.array_equal(own2, 1 - new_owner) and eff2 == eff
    st = pd.reduce_stats({"ncells": 10 + rank, "nfe": 100 * (rank + 1), "max_fc": 5 + rank, "min_fc": 3 - rank,
                          "kernel_ms": 1.0 + rank})
    q.put((rank, new_owner.tolist(), eff, st))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_cost_balance_and_redistribution(plm_lib):
    ws = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in range(ws)]
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == res[1][1], "ranks disagree on the distribution map"
    assert res[0][2] == res[1][2] and 0.9 <= res[0][2] <= 1.0
    st = res[0][3]
    assert st == res[1][3]
    assert st["ncells"] == 21 and st["nfe"] == 300 and st["max_fc"] == 6 and st["min_fc"] == 2 and st["kernel_ms"] == 2.0


def test_chop_for_chemistry():
    from pelelm_b200 import dist as pd
    boxes = [((0, 0, 0), (63, 63, 63))]
    ba = pd.chop_for_chemistry(boxes, (64, 64, 64), nranks=8, threshold=4)
    assert len(ba) >= 32
    cells = sum((h[0] - l[0] + 1) * (h[1] - l[1] + 1) * (h[2] - l[2] + 1) for l, h in ba)
    assert cells == 64 ** 3
    assert all(min(h[d] - l[d] + 1 for d in range(3)) >= 16 for l, h in ba)
    # never below 16 a side even if the threshold cannot be met
    ba = pd.chop_for_chemistry([((0, 0, 0), (31, 31, 31))], (32, 32, 32), nranks=8, threshold=100)
    assert all(min(h[d] - l[d] + 1 for d in range(3)) >= 8 for l, h in ba) and len(ba) <= 64
    # already enough boxes: untouched
    many = [((i * 16, 0, 0), (i * 16 + 15, 15, 15)) for i in range(40)]
    assert pd.chop_for_chemistry(many, (16, 16, 16), nranks=8, threshold=4) == many


def test_max_size_vs_reference():
    """dist.max_size against the pieces AMReX's own BoxList::maxSize produced (tests/golden/maxsize.npz, written by
    make_golden.py through oracle/ref_knapsack.cpp::ref_max_size): same pieces, same order -- non-divisible sides, sides that
    share a power of two with the chunk, boxes with a non-zero origin, directions that need no cut."""
    from pelelm_b200 import dist as pd
    g = np.load(os.path.join(ROOT, "tests", "golden", "maxsize.npz"))
    i = 0
    while "lo%d" % i in g:
        lo, hi, ch = (tuple(int(v) for v in g[k + str(i)]) for k in ("lo", "hi", "chunk"))
        got = pd.max_size([(lo, hi)], ch)
        want = [(tuple(int(v) for v in r[:3]), tuple(int(v) for v in r[3:])) for r in g["pieces%d" % i]]
        assert got == want, (i, lo, hi, ch)
        i += 1
    assert i == 8
