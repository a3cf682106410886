"""Multi-GPU plumbing of the chemistry advance: one process per GPU, torch.distributed (NCCL on the GPU box,
gloo in the CPU tests).  Chemistry has no halo, so the only collectives are the ones the reference has
(SURVEY.md section 8e):

  * per-box cost vector: all_gather of each rank's box costs (replaces MPI_Gatherv + Bcast,
    PeleLM::getFuncCountDM, Source/PeleLM.cpp:5582-5615), then every rank runs the same deterministic knapsack
    (DistributionMapping::KnapSackProcessorMap, AMReX_DistributionMapping.cpp:846-870 -> plm_knapsack);
  * statistics: all_reduce sum/max/min (ParallelDescriptor::ReduceRealMin/Max, PeleLM.cpp:5838-5839).

Box payload motion when the chemistry DM differs from the state DM (FabArray::ParallelCopy, PeleLM.cpp:5695-5696,
:5780) is `redistribute`: grouped point-to-point sends of whole boxes.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import capi


def max_size(boxes, chunk):
    """BoxList::maxSize(chunk) (AMReX_BoxList.cpp:774-827): chop every box so that no side is longer than chunk[d].
    The common power of two of the side and the chunk is kept as the blocking factor of the pieces (a 48-cell side chopped
    at 32 gives 32 + 16, not 24 + 24), the longer pieces come first, and the pieces are emitted x fastest."""
    out = []
    for lo, hi in boxes:
        cuts = []
        for d in range(3):
            n = hi[d] - lo[d] + 1
            c = max(1, int(chunk[d]))
            if n <= c:
                cuts.append([(lo[d], hi[d])])
                continue
            ratio, bs, nlen = 1, c, n
            while bs % 2 == 0 and nlen % 2 == 0:
                ratio, bs, nlen = ratio * 2, bs // 2, nlen // 2
            k = (nlen + bs - 1) // bs
            sz = nlen // k
            extra = nlen - sz * k
            edges = []
            for i in range(k):
                if i < extra:
                    a = i * (sz + 1) * ratio
                    b = a + (sz + 1) * ratio - 1
                else:
                    a = (i * sz + extra) * ratio
                    b = a + sz * ratio - 1
                edges.append((lo[d] + a, lo[d] + b))
            cuts.append(edges)
        for kz in cuts[2]:
            for jy in cuts[1]:
                for ix in cuts[0]:
                    out.append(((ix[0], jy[0], kz[0]), (ix[1], jy[1], kz[1])))
    return out


def chop_for_chemistry(boxes, max_grid_size, nranks, threshold=4, min_side=16):
    """Box chopping heuristic of advance_chemistry (PeleLM.cpp:5660-5687): halve the chunk per direction
    (z first) until there are >= threshold*nranks boxes, never going below min_side cells a side."""
    target = threshold * nranks
    ba = list(boxes)
    done = len(ba) >= target
    cnt = 1
    while not done:
        chunk = [g // cnt for g in max_grid_size]
        if any(c < min_side for c in chunk):
            break
        for j in (2, 1, 0):
            if len(ba) >= target:
                break
            chunk[j] //= 2
            ba = max_size(ba, chunk)
            if len(ba) >= target:
                done = True
        cnt *= 2
    return ba


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def gather_box_costs(local_costs, owner_of_box, group=None):
    """All ranks get the full cost vector in box order. local_costs: costs of the boxes this rank owns, in
    ascending box id; owner_of_box: current rank of every box (same on all ranks)."""
    owner = np.asarray(owner_of_box)
    nb = owner.size
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        full = np.zeros(nb, dtype=np.int64)
        full[owner == 0] = np.asarray(local_costs, dtype=np.int64)
        return full
    ws, rk = dist.get_world_size(group), dist.get_rank(group)
    dev = _device()
    # zero-padded vector + sum == gather (each box has exactly one owner)
    v = torch.zeros(nb, dtype=torch.int64, device=dev)
    mine = np.nonzero(owner == rk)[0]
    assert len(mine) == len(local_costs), (len(mine), len(local_costs))
    if len(mine):
        v[torch.as_tensor(mine, device=dev)] = torch.as_tensor(np.asarray(local_costs, dtype=np.int64), device=dev)
    dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
    return v.cpu().numpy()


def func_count_dm(costs, nranks, default_owner=None, rank_order=None):
    """getFuncCountDM's decision: FC all zero -> keep the default mapping (PeleLM.cpp:5545-5549), else knapsack.
    rank_order = least_used_ranks(...) gives the reference's sort = true bucket -> rank assignment; None = identity
    (every rank holds the same amount of state, as in the synthetic configurations)."""
    costs = np.asarray(costs, dtype=np.int64)
    if costs.size == 0 or costs.max() <= 0:
        if default_owner is not None:
            return np.asarray(default_owner, dtype=np.int32), 1.0
        return (np.arange(costs.size) % nranks).astype(np.int32), 1.0
    return capi.knapsack(costs, nranks, rank_order=rank_order)


def least_used_ranks(local_kbytes, group=None):
    """DistributionMapping::LeastUsedCPUs (AMReX_DistributionMapping.cpp:191-235): one all-gather of the kilobytes each rank
    holds (the reference counts FAB bytes; a GPU rank passes torch.cuda.memory_allocated() // 1024), then the ranks by
    ascending load, equal on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return np.zeros(1, dtype=np.int32)
    ws = dist.get_world_size(group)
    dev = _device()
    mine = torch.tensor([int(local_kbytes)], dtype=torch.int64, device=dev)
    allv = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(ws)]
    dist.all_gather(allv, mine, group=group)
    return capi.least_used_ranks([int(v.item()) for v in allv])


def reduce_stats(stats: dict, group=None):
    """Sum the additive integrator counters over ranks; max/min for the extrema and the kernel time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(stats)
    dev = _device()
    add = ["ncells", "nskipped", "nst", "nfe", "nje", "nsetups", "nni", "ncfn", "netf", "nfailed"]
    s = torch.tensor([float(stats.get(k, 0)) for k in add], dtype=torch.float64, device=dev)
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    mx = torch.tensor([float(stats.get("max_fc", 0)), float(stats.get("kernel_ms", 0.0)), -float(stats.get("min_fc", 0))],
                      dtype=torch.float64, device=dev)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    out = {k: int(v) for k, v in zip(add, s.tolist())}
    out["max_fc"], out["kernel_ms"], out["min_fc"] = int(mx[0].item()), mx[1].item(), int(-mx[2].item())
    return out


def redistribute(payload: dict, old_owner, new_owner, group=None):
    """ParallelCopy of whole boxes between distribution maps: payload = {box id: tensor} held by this rank under
    old_owner; returns the dict this rank holds under new_owner. All boxes must have the same shape/dtype per id
    on sender and receiver (the receiver allocates from `like`)."""
    old_owner, new_owner = np.asarray(old_owner), np.asarray(new_owner)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(payload)
    rk = dist.get_rank(group)
    out, ops, keep = {}, [], []
    shapes = {b: (tuple(t.shape), t.dtype, t.device) for b, t in payload.items()}
    # shapes of boxes we will receive: exchanged as a small object gather (host side, tiny)
    all_shapes = [None] * dist.get_world_size(group)
    dist.all_gather_object(all_shapes, {b: (s[0], str(s[1])) for b, s in shapes.items()}, group=group)
    dev = _device()
    for b in range(old_owner.size):
        src, dst = int(old_owner[b]), int(new_owner[b])
        if src == dst:
            if src == rk:
                out[b] = payload[b]
            continue
        if src == rk:
            t = payload[b].contiguous()
            keep.append(t)
            ops.append(dist.P2POp(dist.isend, t, dst, group=group, tag=b))
        elif dst == rk:
            shp, dt = all_shapes[src][b]
            t = torch.empty(shp, dtype=getattr(torch, dt.split(".")[-1]), device=dev)
            out[b] = t
            ops.append(dist.P2POp(dist.irecv, t, src, group=group, tag=b))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out


def move_boxes(boxes: dict, keys, old_owner, new_owner, device=None, group=None):
    """Re-home whole boxes after a new distribution map (advance_chemistry's ParallelCopy into the chemistry MultiFabs,
    PeleLM.cpp:5692-5697): `boxes` = {box id: {key: tensor}} of the boxes this rank owns under old_owner; every tensor
    named in `keys` travels; returns the dict of the boxes owned under new_owner (tensors of received boxes live on
    `device`, kept boxes are the original objects)."""
    old_owner, new_owner = np.asarray(old_owner), np.asarray(new_owner)
    rank = dist.get_rank(group) if (dist.is_available() and dist.is_initialized()) else 0
    mine = [b for b in boxes if old_owner[b] == rank]
    out = {b: dict(boxes[b]) for b in mine if new_owner[b] == rank}
    for key in keys:
        payload = {b: (boxes[b][key].to(device) if device is not None else boxes[b][key]) for b in mine}
        moved = redistribute(payload, old_owner, new_owner, group=group)
        for b, t in moved.items():
            if old_owner[b] != rank:
                out.setdefault(b, {"id": b})[key] = t
    return out
