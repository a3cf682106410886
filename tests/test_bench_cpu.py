"""Host logic of bench.py that runs without a GPU: the time budget of the nested workloads (every rank must take rank 0's
decision, because the workloads contain collectives) and the workload bookkeeping."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_budget_single_rank():
    import bench
    assert bench.elapsed_s() > 0.0
    assert bench.fits_budget(None, 100.0, 760.0, now_s=600.0)
    assert not bench.fits_budget(None, 100.0, 760.0, now_s=700.0)
    assert bench.fits_budget({"world": 1}, 135.0, 800.0, now_s=665.0)
    # the driver's scaling run: 870 s per N; the budgets leave room for the process to wind down
    assert bench.BUDGET_S <= bench.BUDGET_LAST_S < 870.0 - 60.0
    # every nested workload has a cost estimate, and all of them together fit behind a 20 + 5 step headline (~560 s)
    assert set(bench.NESTED_COST_S) == {"dodecane_lu", "lidryer_rk64", "tripleflame", "dodecane_lu_256cubed"}
    assert 600.0 + sum(v for k, v in bench.NESTED_COST_S.items() if k != "dodecane_lu_256cubed") <= bench.BUDGET_S
    # N = 8: no small dodecane entry, no CPU baselines (~575 s headline), then the 256^3 pass
    assert 575.0 + bench.NESTED_COST_S["lidryer_rk64"] + bench.NESTED_COST_S["tripleflame"] + bench.NESTED_COST_S["dodecane_lu_256cubed"] \
        <= bench.BUDGET_LAST_S + 10.0


def _worker(rank, ws, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    import bench
    ctx = {"world": ws, "rank": rank, "dev": torch.device("cpu")}
    # rank 0 has time left, rank 1 (a slower start) would say no: both run the workload
    a = bench.fits_budget(ctx, 100.0, 760.0, now_s=600.0 if rank == 0 else 750.0)
    # rank 0 is out of time, rank 1 is not: both skip
    b = bench.fits_budget(ctx, 100.0, 760.0, now_s=700.0 if rank == 0 else 10.0)
    q.put((rank, a, b))
    dist.barrier()
    dist.destroy_process_group()


def test_budget_decision_is_rank_zeros():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert out == [(0, True, False), (1, True, False)]
