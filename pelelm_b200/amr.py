"""Per-level chemistry advance over an AMR hierarchy with cost-weighted box distribution (BASELINE.json configs[4],
SURVEY.md section 8d "cfg5").

What the reference does per level (PeleLM::advance_chemistry, Source/PeleLM.cpp:5625-5844), restated for one process per GPU:

  1. `dm = getFuncCountDM(ba, ngrow)` (:5529-5623, :5689): per-box cost = sum of the previous pass's FC, gathered on every
     rank, knapsack -> chemistry owner of every box (all-zero FC keeps the state's own map);
  2. `STemp / FTemp.ParallelCopy` (:5692-5696): state and forcing of every box travel to its chemistry owner;
  3. the MFIter loop (:5725-5775): MKS->CGS, react, CGS->MKS per box -- here ONE fused launch over every local box of the
     level (`plm_react_boxes`; `plm_advance_chemistry_box` per box when `multi=False`), which also forms
     I_R = (new - old)/dt - F (:5789-5803);
  4. `mf_new / FC.ParallelCopy` back to the state's map (:5780, :5812);
  5. `average_down(fine_React, React_new)` (:5818-5822) when `avg_down_chem` and a finer level with old I_R exists: the I_R
     of covered coarse cells is replaced by the average of the fine level's previous I_R.

Every level is advanced over all of its cells (the reference's react_mask is 1 everywhere without EB, :5706-5711).
The hierarchy is the synthetic "TripleFlame" of SURVEY.md section 8d: mixture-fraction layer z(x) = (1 + tanh((x - 0.6 L)/(0.05 L)))/2
(Docs/source/manual/Tutorials_TripleFlame.rst:164) plus a hot Gaussian kernel on the layer; levels refined x2 around both;
dt_L = dt / 2^L (ReactEval's convention, PelePhysics Testing/Exec/ReactEval/main.cpp:524).

The integration itself is passed in as `advance_box` (on the GPU: `Reactor.advance_chemistry`), so that the host logic --
costs, knapsack, box motion, average-down -- runs under gloo on CPU in tests/test_amr_cpu.py.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch
import torch.distributed as dist

from . import dist as pd
from . import synth


@dataclass
class Level:
    lev: int
    domain: tuple            # cells per direction of this level's index space
    boxes: list              # [(lo, hi)] inclusive corners in this level's index space
    owner: np.ndarray        # state DistributionMapping: rank of every box
    data: dict = field(default_factory=dict)   # {box id: {"S": [ns+2][nc] rhoY, rhoH, T (MKS), "F": [ns+1][nc], "I_R": [ns][nc],
    #                                                       "I_R_old": [ns][nc], "FC": [nc], "mask": [nc] int32}} of the local boxes

    def shape(self, b):
        lo, hi = self.boxes[b]
        return tuple(hi[d] - lo[d] + 1 for d in range(3))

    def ncells(self, b=None):
        if b is not None:
            return int(np.prod(self.shape(b)))
        return sum(int(np.prod(self.shape(i))) for i in range(len(self.boxes)))


# ---------------------------------------------------------------------------------------------------------------------
# synthetic TripleFlame hierarchy
# ---------------------------------------------------------------------------------------------------------------------
KERNEL_SIGMA = 0.08   # of Ly
KERNEL_AT = (0.6, 0.25, 0.5)


def _fields(xc, yc, zc):
    """Mixture-fraction layer and hot kernel at normalised coordinates in [0,1]^3 (y is the long, streamwise axis)."""
    Z = 0.5 * (1.0 + np.tanh((xc - 0.6) / 0.05))
    # the domain is 1 : 2 : 1, distances in units of Ly
    r2 = (0.5 * (xc - KERNEL_AT[0])) ** 2 + (yc - KERNEL_AT[1]) ** 2 + (0.5 * (zc - KERNEL_AT[2])) ** 2
    g = np.exp(-r2 / (2.0 * KERNEL_SIGMA ** 2))
    return Z, g, r2


def _tagged(lev, xc, yc, zc):
    """Refinement criterion producing level lev+1: the mixing layer and the kernel, nested with level."""
    _, _, r2 = _fields(xc, yc, zc)
    w = 0.10 / (2 ** lev)
    return (np.abs(xc - 0.6) < w) | (r2 < ((3.0 - lev) * KERNEL_SIGMA) ** 2)


def tripleflame_boxes(n0=(32, 64, 32), nlev=3, box=32):
    """Box lists of an nlev-level hierarchy: level 0 covers the domain in box^3 grids; a box^3 grid of level l+1 (aligned on
    box, so that its coarsened image lies inside one coarse grid) exists wherever a cell of its footprint is tagged."""
    levels = []
    for lev in range(nlev):
        dom = tuple(n * 2 ** lev for n in n0)
        nb = [max(1, d // box) for d in dom]
        boxes = []
        for kz in range(nb[2]):
            for jy in range(nb[1]):
                for ix in range(nb[0]):
                    lo = (ix * box, jy * box, kz * box)
                    hi = tuple(min(lo[d] + box, dom[d]) - 1 for d in range(3))
                    if lev > 0:
                        ax = [(np.arange(lo[d], hi[d] + 1) + 0.5) / dom[d] for d in range(3)]
                        X, Y, Zc = np.meshgrid(ax[0], ax[1], ax[2], indexing="ij")
                        if not _tagged(lev - 1, X, Y, Zc).any():
                            continue
                    boxes.append((lo, hi))
        levels.append((dom, boxes))
    return levels


def tripleflame_state(mech_name, dom, lo, hi, dt=None, forcing=False):
    """State of one box in PeleLM's units (MKS), component-major with x fastest (FAB layout):
    S [ns+2][nc] = rhoY, rhoH, T; F [ns+1][nc] forcing (zero unless forcing: a smooth 1 % / dt source)."""
    m = synth.mech(mech_name)
    ns = m.ns
    ax = [(np.arange(lo[d], hi[d] + 1) + 0.5) / dom[d] for d in range(3)]
    Zc, Yc, Xc = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")   # C order: z slowest, x fastest
    xc, yc, zc = Xc.ravel(), Yc.ravel(), Zc.ravel()
    Zmix, g, _ = _fields(xc, yc, zc)
    phi = 0.5 * np.exp(np.log(4.0) * Zmix)       # lean 0.5 ... rich 2 across the layer
    T = 900.0 + 1500.0 * g                       # hot kernel on a 900 K stream
    c = g
    urad = np.stack([np.full_like(T, 0.5)] * 4)
    Y, rho, h = synth.flame_state(m, mech_name, T, phi, c, urad)
    S = np.empty((ns + 2, T.size))
    S[:ns] = (Y * rho[:, None]).T * 1.0e3        # g/cm^3 -> kg/m^3
    S[ns] = rho * h * 0.1                        # erg/cm^3 -> J/m^3
    S[ns + 1] = T
    F = np.zeros((ns + 1, T.size))
    if forcing:
        assert dt is not None
        w = np.sin(2.0 * np.pi * yc) * np.cos(2.0 * np.pi * zc)
        s = 1e-2 * w[None, :] * S[:ns] / dt * np.cos(np.arange(ns))[:, None]
        s -= S[:ns] / S[:ns].sum(0, keepdims=True) * s.sum(0, keepdims=True)
        F[:ns] = s
        F[ns] = 1e-2 * w * np.abs(S[ns]) / dt
    return S, F


def build_tripleflame(mech_name, nranks, rank, n0=(32, 64, 32), nlev=3, box=32, device="cpu", forcing=False, dt=1e-5):
    """The hierarchy with this rank's boxes filled; state DM = round robin (the default map getFuncCountDM starts from,
    PeleLM.cpp:5537-5538); FC = 1 per cell as PeleLM initialises it (PeleLM.cpp:2006)."""
    ns = synth.mech(mech_name).ns
    levels = []
    for lev, (dom, boxes) in enumerate(tripleflame_boxes(n0, nlev, box)):
        L = Level(lev, dom, boxes, (np.arange(len(boxes)) % nranks).astype(np.int32))
        for b in np.nonzero(L.owner == rank)[0]:
            lo, hi = boxes[b]
            S, F = tripleflame_state(mech_name, dom, lo, hi, dt=dt / 2 ** lev, forcing=forcing)
            nc = S.shape[1]
            L.data[int(b)] = {
                "id": int(b), "S": torch.from_numpy(S).to(device), "F": torch.from_numpy(F).to(device),
                "I_R": torch.zeros((ns, nc), dtype=torch.float64, device=device),
                "I_R_old": torch.zeros((ns, nc), dtype=torch.float64, device=device),
                "FC": torch.ones(nc, dtype=torch.float64, device=device),
                "mask": torch.ones(nc, dtype=torch.int32, device=device)}
        levels.append(L)
    return levels


# ---------------------------------------------------------------------------------------------------------------------
# advance
# ---------------------------------------------------------------------------------------------------------------------
def _rank(group=None):
    return dist.get_rank(group) if (dist.is_available() and dist.is_initialized()) else 0


def _world(group=None):
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


def local_box_costs(L: Level, cost_fn=None):
    """Sum of FC per local box, ascending box id (the `wgts` of getFuncCountDM, PeleLM.cpp:5556-5580).
    cost_fn(FCall, offsets) -> int64 tensor is the library reduction on the GPU (capi.box_costs)."""
    ids = sorted(L.data)
    if not ids:
        return np.zeros(0, dtype=np.int64)
    if cost_fn is None:
        return np.array([int(torch.sum(L.data[b]["FC"]).item()) for b in ids], dtype=np.int64)
    fc = torch.cat([L.data[b]["FC"] for b in ids])
    off = torch.tensor(np.concatenate([[0], np.cumsum([L.data[b]["FC"].numel() for b in ids])]), dtype=torch.int64,
                       device=fc.device)
    return cost_fn(fc, off).cpu().numpy().astype(np.int64)


def average_down(L: Level, fine: Level, key_fine="I_R_old", key_crse="I_R", ratio=2, group=None):
    """amrex::average_down for cell-centred data (AMReX_MultiFabUtil.cpp: plain mean over the ratio^3 fine cells) of
    fine[key_fine] onto the covered cells of L[key_crse]; patches travel from the fine box's owner to the coarse box's."""
    rk = _rank(group)
    # covering coarse box of every fine box (fine grids are aligned so that one exists)
    cover = np.full(len(fine.boxes), -1, dtype=np.int64)
    clo = np.array([b[0] for b in L.boxes])
    chi = np.array([b[1] for b in L.boxes])
    for f, (lo, hi) in enumerate(fine.boxes):
        l = np.array(lo) // ratio
        h = np.array(hi) // ratio
        hit = np.nonzero(np.all(clo <= l, 1) & np.all(chi >= h, 1))[0]
        assert hit.size == 1, "fine grid %d is not nested in one coarse grid" % f
        cover[f] = hit[0]
    patches = {}
    for f, bx in fine.data.items():
        nx, ny, nz = fine.shape(f)
        v = bx[key_fine].reshape(-1, nz // ratio, ratio, ny // ratio, ratio, nx // ratio, ratio)
        patches[f] = v.mean(dim=(2, 4, 6)).contiguous()
    dest = L.owner[cover].astype(np.int32)
    got = pd.redistribute(patches, fine.owner, dest, group=group)
    for f, p in got.items():
        if dest[f] != rk:
            continue
        c = int(cover[f])
        nx, ny, nz = L.shape(c)
        lo = [fine.boxes[f][0][d] // ratio - L.boxes[c][0][d] for d in range(3)]
        px, py, pz = p.shape[3], p.shape[2], p.shape[1]
        tgt = L.data[c][key_crse].reshape(-1, nz, ny, nx)
        tgt[:, lo[2]:lo[2] + pz, lo[1]:lo[1] + py, lo[0]:lo[0] + px] = p.to(tgt.device)


def advance_chemistry_level(L: Level, dt, advance_box, fine: Level | None = None, avg_down_chem=True, use_knapsack=True,
                            cost_fn=None, group=None, device=None, wait=None, rank_order=None):
    """One PeleLM::advance_chemistry of level L (steps 1-5 of the module docstring).

    advance_box(box: dict, shape, dt, slot) advances box["S"] in place, writes box["I_R"] and box["FC"]; `wait()` (if
    given) blocks until every advance_box call has finished (asynchronous launches on the GPU); rank_order =
    pelelm_b200.dist.least_used_ranks(...) hands the heaviest bucket to the least-used rank (the reference's sort = true).
    Returns {"chem_owner", "costs", "efficiency", "moved"}."""
    rk, nr = _rank(group), _world(group)
    # the new time level starts: what was new becomes old (the state swap of the time step)
    for bx in L.data.values():
        bx["I_R_old"], bx["I_R"] = bx["I_R"], bx["I_R_old"]
    costs = pd.gather_box_costs(local_box_costs(L, cost_fn), L.owner, group=group)
    if use_knapsack:
        chem_owner, eff = pd.func_count_dm(costs, nr, default_owner=L.owner, rank_order=rank_order)
    else:
        chem_owner, eff = L.owner.copy(), None
    chem_owner = np.asarray(chem_owner, dtype=np.int32)
    moved = int(np.count_nonzero(chem_owner != L.owner))
    # STemp / FTemp in the chemistry map; I_R and FC are produced there
    work = pd.move_boxes(L.data, ("S", "F", "mask"), L.owner, chem_owner, device=device, group=group) if moved else L.data
    todo = []
    for i, b in enumerate(sorted(work)):
        bx = work[b]
        if "I_R" not in bx:   # a received box: outputs are allocated at the chemistry owner
            ns = bx["F"].shape[0] - 1
            bx["I_R"] = torch.empty((ns, bx["S"].shape[1]), dtype=torch.float64, device=bx["S"].device)
            bx["FC"] = torch.empty(bx["S"].shape[1], dtype=torch.float64, device=bx["S"].device)
        todo.append((bx, L.shape(b)))
    if hasattr(advance_box, "all_boxes"):   # one launch over every local box of the level (plm_react_boxes)
        advance_box.all_boxes(todo, dt)
    else:
        for i, (bx, shape) in enumerate(todo):
            advance_box(bx, shape, dt, i)
    if wait is not None:
        wait()
    if moved:
        back = pd.move_boxes(work, ("S", "I_R", "FC"), chem_owner, L.owner, device=device, group=group)
        for b, bx in back.items():
            if chem_owner[b] != rk:   # came home: results into the state's own arrays
                for key in ("S", "I_R", "FC"):
                    L.data[b][key].copy_(bx[key])
    if avg_down_chem and fine is not None:
        average_down(L, fine, "I_R_old", "I_R", group=group)
    return {"chem_owner": chem_owner, "costs": costs, "efficiency": eff, "moved": moved}


def advance_hierarchy(levels, dt, advance_box, **kw):
    """One coarse step: every level once with dt / 2^lev, coarse level first as Amr::timeStep walks them, so that the
    coarse average-down reads the fine level's OLD I_R before that level swaps its own time levels (advance_setup,
    IAMR NavierStokesBase.cpp:712-723, is per level) -- literally what get_old_data(RhoYdot_Type) of level+1 holds then."""
    out = []
    for L in levels:
        fine = levels[L.lev + 1] if L.lev + 1 < len(levels) else None
        out.append(advance_chemistry_level(L, dt / 2 ** L.lev, advance_box, fine=fine, **kw))
    return out


def gpu_advance(reactor, streams=None, multi=True):
    """(advance_box, wait) for advance_chemistry_level on the GPU: one fused launch over every local box of the level
    (`multi`, `plm_react_boxes`), or one `plm_advance_chemistry_box` launch per box spread round robin over `streams`
    (device arrays, the box as its own FAB); wait() drains them and returns the integrator statistics accumulated since
    the previous wait()."""
    ns = reactor.ns
    launches = [0]

    def advance_box(bx, shape, dt, slot):
        S, F = bx["S"], bx["F"]
        st = streams[slot % len(streams)] if streams else None
        lo, hi = [0, 0, 0], [shape[0] - 1, shape[1] - 1, shape[2] - 1]
        if st is not None:
            st.wait_stream(torch.cuda.current_stream())
        reactor.advance_chemistry(lo, hi, S[:ns], S[ns], S[ns + 1], F[:ns], F[ns], bx["I_R"], bx["FC"], bx["mask"], dt,
                                  stream=st.cuda_stream if st is not None else None, sync=False)
        launches[0] += 1

    def all_boxes(todo, dt):
        """every local box of the level in one call: with the lane = cell kernels one launch whose work queue runs over the
        cells of all boxes (the queue's tail is paid once per level, not once per box); a loop of launches otherwise"""
        if not todo:
            return
        boxes = []
        for bx, shape in todo:
            S, F = bx["S"], bx["F"]
            boxes.append({"lo": [0, 0, 0], "hi": [shape[0] - 1, shape[1] - 1, shape[2] - 1], "rY": S[:ns], "rEner": S[ns],
                          "T": S[ns + 1], "rYsrc": F[:ns], "rEner_src": F[ns], "I_R": bx["I_R"], "FC": bx["FC"], "mask": bx["mask"]})
        reactor.react_boxes(boxes, dt, mks=True, stream=torch.cuda.current_stream().cuda_stream, sync=False)
        launches[0] += len(boxes) if reactor.path_is_per_box() else 1

    def wait():
        out = reactor.stats()   # waits for every outstanding launch of this reactor
        if out.get("nfailed", 0):   # the reference aborts inside CVODE (cvodeErrHandler); do not carry a failed cell's state on
            raise RuntimeError("chemistry integration failed in %d cells (status per cell: Reactor.cell_status)" % out["nfailed"])
        if streams:
            for st in streams:
                torch.cuda.current_stream().wait_stream(st)
        return out

    advance_box.launches = launches
    if multi:
        advance_box.all_boxes = all_boxes
    return advance_box, wait
