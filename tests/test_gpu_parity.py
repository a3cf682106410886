"""Parity of the CUDA path (through the C ABI of libplmchem.so) with the checkers, on a B200.

Checkers: (1) the committed golden fixtures = outputs of the UNMODIFIED reference, (2) the oracle port run live
on seeded inputs.  Bar (BASELINE.json north_star): final Y_k and T relative error <= 1e-6 at rtol = atol = 1e-10,
identical mask/skip behaviour; rhoH bit-exact; FC identical where the port's is.
"""
import numpy as np
import pytest

import os

from conftest import MECHS, GOLDEN, golden, golden_plm, golden_e, check_state, rel_errors
from golden.make_golden import CASES, DT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi(plm_lib):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from pelelm_b200 import capi as c
    return c


# kernel families to hold to the fixtures: "auto" = what a caller gets (lane = cell kernels for LiDryer / drm19 where they cover
# the request, warp-per-cell otherwise), "warp" = the warp-per-cell kernels forced (they stay the path of dodecane_lu, of the
# difference-quotient Jacobian and of the internal-energy BDF, and must stay right for the small mechanisms too)
def _paths(mech):
    return ("auto", "warp") if mech != "dodecane_lu" else ("auto",)


def _gpu_react(capi, mech, integ, ikw, b, typ=None, box=None, path="auto"):
    kw = dict(ikw)
    rk = integ == "ReactorRK64"
    R = capi.Reactor(mech, integ).init(rtol=kw.get("rtol", kw["atol"]), atol=kw["atol"], max_order=kw.get("max_order", 2),
                                       clean_init_massfrac=kw.get("clean_init_massfrac", 0),
                                       solve_type=kw.get("solve_type", "denseAJ_direct")).set_path(path)
    if typ is not None:
        R.set_typ_vals_ode(typ)
    nc = b["T"].size
    a = [b["rY"].copy(), b["rYsrc"].copy(), b["T"].copy(), b["rE"].copy(), b["rEsrc"].copy(), np.zeros(nc)]
    st = R.react([0, 0, 0], [nc - 1, 0, 0], a[0], a[1], a[2], a[3], a[4], a[5], b["mask"].copy(), DT)
    status = R.cell_status(nc)
    R.close()
    return (a[0], a[2], a[3], a[5]), st, status


@pytest.mark.parametrize("path", ["auto", "warp"])
@pytest.mark.parametrize("mech", MECHS)
@pytest.mark.parametrize("case", list(CASES) + ["reval"])
def test_react_vs_reference_golden(capi, mech, case, path):
    g = golden(mech)
    if case + "/rY" not in g:
        pytest.skip("case not generated for this mechanism")
    if path not in _paths(mech):
        pytest.skip("one kernel family for this mechanism")
    integ, ikw, _ = CASES.get(case, ("ReactorCvode", dict(rtol=1e-10, atol=1e-10, max_order=2), None))
    b = {k: g[case + "/in_" + k] for k in ("rY", "rYsrc", "T", "rE", "rEsrc", "mask")}
    typ = g[case + "/typ"] if case + "/typ" in g else None
    got, st, status = _gpu_react(capi, mech, integ, ikw, b, typ, path=path)
    atol = ikw["atol"] * (typ[:-1, None] if typ is not None else 1.0)
    ref = (g[case + "/rY"], g[case + "/T"], g[case + "/rE"], g[case + "/FC"])
    small = mech != "dodecane_lu"
    exact_fc = small and case != "cv_dq"
    cmp_rtol = max(1e-6, 100 * ikw.get("rtol", 0.0))
    if case == "cv_dq":  # the difference-quotient Jacobian amplifies last-bit RHS noise into the Newton iteration
        cmp_rtol = 1e-5
    check_state(got, ref, atol, tag="%s/%s" % (mech, case), rtol=cmp_rtol, fc_equal_min=0.9 if exact_fc else None,
                fc_rel=0.05 if small else 0.15)
    assert not status.any()
    m = b["mask"] == -1
    if integ == "ReactorCvode":
        assert st["nskipped"] == int(m.sum()) and st["ncells"] == int((~m).sum())
        if m.any():
            assert np.array_equal(got[0][:, m], b["rY"][:, m]) and np.array_equal(got[1][m], b["T"][m])
            assert np.all(got[3][m] == 0)
    else:
        assert st["nskipped"] == 0  # ReactorRK64 ignores the mask (ReactorRK64.cpp:176)


@pytest.mark.parametrize("mech,nc", [("LiDryer", 512), ("drm19", 512), ("dodecane_lu", 48)])
def test_react_vs_port_live(capi, port_lib, mech, nc):
    from pelelm_b200 import synth
    b = synth.synthetic_batch(mech, nc, seed=4242, forcing=True, mask_every=97)
    ikw = dict(rtol=1e-10, atol=1e-10, max_order=2)
    P = port_lib.Port(mech).init(**ikw)
    ref = P.react_soa(b["rY"], b["rYsrc"], b["T"], b["rE"], b["rEsrc"], b["mask"], DT, nthreads=8)
    got, st, status = _gpu_react(capi, mech, "ReactorCvode", ikw, b)
    small = mech != "dodecane_lu"
    check_state(got, ref, 1e-10, tag=mech, fc_equal_min=0.95 if small else None, fc_rel=0.02 if small else 0.1)
    assert not status.any()


@pytest.mark.parametrize("mech", MECHS)
def test_rhs_vs_reference_and_port(capi, port_lib, mech):
    g = golden(mech)
    R = capi.Reactor(mech, "ReactorCvode").init()
    f = R.eval_rhs(g["rhs/y"], g["rhs/rhoe"], g["rhs/rhoesrc"], g["rhs/rYsrc"], float(g["rhs/t"]))
    ref = g["rhs/ydot"]
    err = np.abs(f - ref) / np.abs(ref).max(1, keepdims=True)
    assert err.max() < 1e-9, err.max()
    fp = port_lib.Port(mech).rhs(g["rhs/y"], g["rhs/rhoe"], g["rhs/rhoesrc"], g["rhs/rYsrc"], float(g["rhs/t"]))
    err = np.abs(f - fp) / np.abs(fp).max(1, keepdims=True)
    assert err.max() < 1e-11, err.max()
    R.close()


@pytest.mark.parametrize("mech", MECHS)
def test_jac_vs_reference(capi, mech):
    g = golden(mech)
    R = capi.Reactor(mech, "ReactorCvode").init()
    J = R.eval_jac(g["jac/y"])
    ref = g["jac/J"]
    err = np.abs(J - ref) / np.abs(ref).max(1, keepdims=True)
    assert err.max() < 1e-11, err.max()
    R.close()


@pytest.mark.parametrize("mech", MECHS)
def test_big_batch_vs_reference(capi, mech):
    """The headline configuration (ReactorCvode, rtol = atol = 1e-10, order <= 2, dt = 1e-5) on every cell of
    tests/golden/<mech>_big.npz: 1024 / 1024 / 512 cells of the bench's own synthetic states advanced by the unmodified reference.
    Prints the measured relative errors."""
    from pelelm_b200 import synth
    g = np.load(os.path.join(GOLDEN, mech + "_big.npz"))
    n, ch = g["T"].size, int(g["chunk"])
    parts = [synth.synthetic_batch(mech, ch, cell0=c, seed=synth.DEFAULT_SEED + 301, dt=DT, forcing=True, mask_every=61)
             for c in range(0, n, ch)]
    b = {k: np.ascontiguousarray(np.concatenate([p[k] for p in parts], axis=-1)) for k in parts[0]}
    assert np.array_equal(b["T"], g["in_T"])
    ref = (g["rY"], g["T"], g["rE"], g["FC"])
    small = mech != "dodecane_lu"
    for path in _paths(mech):
        got, st, status = _gpu_react(capi, mech, "ReactorCvode", dict(rtol=1e-10, atol=1e-10, max_order=2), b, path=path)
        check_state(got, ref, 1e-10, tag="%s/big/%s" % (mech, path), fc_equal_min=0.98 if small else 0.5,
                    fc_rel=0.01 if small else 0.05)
        assert not status.any()
        r1, r2 = rel_errors(got[0], ref[0])
        same = got[3] == ref[3]   # cells that took the reference's step sequence: there the difference is rounding only
        s1, s2 = rel_errors(got[0][:, same], ref[0][:, same])
        print("%s [%s kernels], %d cells: max relative error of rhoY_k over Y_k > 1e-12: %.2e; over rhoY_k > atol: %.2e; T: %.2e; "
              "FC identical on %.1f%% of cells (there: %.2e / %.2e), total FC off by %.2e" %
              (mech, path, n, r1, r2, np.max(np.abs(got[1] / ref[1] - 1)), 100 * np.mean(same), s1, s2,
               abs(got[3].sum() / ref[3].sum() - 1)))
        # dodecane_lu: ~2500 steps per cell and a Gauss-Jordan inverse instead of LU: rounding differences accumulate to 1e-5
        assert (s2 <= 1e-6 and r2 <= 1e-4) if small else (s2 <= 1e-4 and r2 <= 1e-3)


@pytest.mark.parametrize("mech", MECHS)
def test_flat_api_vs_reference(capi, mech):
    """react(rY_in[ncells][N+1], rYsrc_in, rX_in, rX_src_in, dt, time, ncells) (ReactorBase.H:43-55) against the reference's own
    1dArray react called with one cell per call (tests/golden/<mech>_flat.npz); every cell is its own CVODE problem here
    whatever ncells is (include/plm_chem.h).  Includes ReactorRK64's 1dArray quirk: no step-size clamp inside the loop."""
    g = np.load(os.path.join(GOLDEN, mech + "_flat.npz"))
    for name, integ in (("cv", "ReactorCvode"), ("rk", "ReactorRK64")):
        if name + "/y" not in g:
            continue
        for path in _paths(mech):
            R = capi.Reactor(mech, integ).init(rtol=1e-10, atol=1e-10, max_order=2).set_path(path)
            y, src = g[name + "/in_y"].copy(), np.ascontiguousarray(g[name + "/in_src"])
            rX, rXs = g[name + "/in_rX"].copy(), np.ascontiguousarray(g[name + "/in_rXsrc"])
            st = R.react_flat(y, src, rX, rXs, DT)
            R.close()
            ok = ~np.isnan(g[name + "/y"]).any(1)
            assert st["ncells"] == rX.size
            assert np.array_equal(ok, ~np.isnan(y).any(1)), (name, path)
            assert np.array_equal(rX[ok], g[name + "/rX"][ok])
            if not ok.any():
                continue
            e = np.abs(y - g[name + "/y"])[ok] / (1e-6 * np.abs(g[name + "/y"][ok]) + 1e-10)
            assert e.max() <= 1.0, (name, path, e.max())
            if mech != "dodecane_lu":  # total work: FC of the cells as the reference counts them
                assert abs(st["nfe" if name == "cv" else "nst"] / g[name + "/FC"][ok].sum() - 1) < 0.01 or not ok.all()


def test_device_pointers_equal_host_pointers(capi):
    import torch
    from pelelm_b200 import synth
    mech, nc = "drm19", 300
    b = synth.synthetic_batch(mech, nc, seed=3, forcing=True, mask_every=11)
    ikw = dict(rtol=1e-8, atol=1e-10, max_order=3)
    got, st, _ = _gpu_react(capi, mech, "ReactorCvode", ikw, b)
    R = capi.Reactor(mech, "ReactorCvode").init(**ikw)
    dev = [torch.from_numpy(b[k].copy()).cuda() for k in ("rY", "rYsrc", "T", "rE", "rEsrc")]
    FC = torch.zeros(nc, dtype=torch.float64, device="cuda")
    mask = torch.from_numpy(b["mask"]).cuda()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        R.react([0, 0, 0], [nc - 1, 0, 0], dev[0], dev[1], dev[2], dev[3], dev[4], FC, mask, DT, stream=s.cuda_stream)
    s.synchronize()
    R.close()
    assert np.array_equal(dev[0].cpu().numpy(), got[0]) and np.array_equal(dev[2].cpu().numpy(), got[1])
    assert np.array_equal(dev[3].cpu().numpy(), got[2]) and np.array_equal(FC.cpu().numpy(), got[3])


def test_subbox_of_fab_with_ghost_cells(capi):
    """react on a tile box strictly inside the FAB (PeleLM.cpp:5727): cells outside the box are untouched."""
    from pelelm_b200 import synth
    mech = "LiDryer"
    flo, fhi = [-2, -2, -1], [9, 5, 1]          # FAB bounds (12 x 8 x 3 cells)
    lo, hi = [0, 0, 0], [7, 3, 0]               # interior tile (8 x 4 x 1)
    nx, ny, nz = 12, 8, 3
    nf = nx * ny * nz
    b = synth.synthetic_batch(mech, nf, seed=21, forcing=True)
    ikw = dict(rtol=1e-10, atol=1e-10, max_order=2)
    R = capi.Reactor(mech, "ReactorCvode").init(**ikw)
    a = [b["rY"].copy(), b["rYsrc"].copy(), b["T"].copy(), b["rE"].copy(), b["rEsrc"].copy(), np.full(nf, -7.0)]
    mask = np.ones(nf, dtype=np.int32)
    R.react(lo, hi, a[0], a[1], a[2], a[3], a[4], a[5], mask, DT, fab_lo=flo, fab_hi=fhi)
    R.close()
    idx = np.arange(nf).reshape(nz, ny, nx)
    inside = idx[1:2, 2:6, 2:10].ravel()
    outside = np.setdiff1d(np.arange(nf), inside)
    assert np.array_equal(a[0][:, outside], b["rY"][:, outside]) and np.all(a[5][outside] == -7.0)
    assert np.all(a[5][inside] > 0)
    sub = {k: (b[k][:, inside].copy() if b[k].ndim == 2 else b[k][inside].copy()) for k in ("rY", "rYsrc", "T", "rE", "rEsrc")}
    sub["mask"] = np.ones(inside.size, dtype=np.int32)
    got, _, _ = _gpu_react(capi, mech, "ReactorCvode", ikw, sub)
    assert np.array_equal(got[0], a[0][:, inside]) and np.array_equal(got[1], a[2][inside])
    assert np.array_equal(got[3], a[5][inside])


def test_edge_sizes(capi):
    from pelelm_b200 import synth
    R = capi.Reactor("LiDryer", "ReactorCvode").init(rtol=1e-8, atol=1e-10)
    z = np.zeros((9, 1))
    st = R.react([0, 0, 0], [-1, 0, 0], z, z, np.zeros(1), np.zeros(1), np.zeros(1), np.zeros(1), None, DT)
    assert st["ncells"] == 0
    for nc in (1, 3, 5, 33):
        b = synth.synthetic_batch("LiDryer", nc, seed=nc)
        FC = np.zeros(nc)
        st = R.react([0, 0, 0], [nc - 1, 0, 0], b["rY"], b["rYsrc"], b["T"], b["rE"], b["rEsrc"], FC, None, DT)
        assert st["ncells"] == nc and np.all(FC > 0)
    R.close()


def test_cell_result_independent_of_batch_position(capi):
    """Size-independent property used at full size: a cell's result depends only on its own inputs
    (bit-exact under any permutation of the batch and any batch size)."""
    from pelelm_b200 import synth
    mech, nc = "drm19", 4096
    b = synth.synthetic_batch(mech, nc, seed=31, forcing=True)
    ikw = dict(rtol=1e-8, atol=1e-10, max_order=2)
    got, _, _ = _gpu_react(capi, mech, "ReactorCvode", ikw, b)
    perm = np.random.default_rng(1).permutation(nc)[: nc // 2 + 7]
    sub = {k: (np.ascontiguousarray(b[k][:, perm]) if b[k].ndim == 2 else b[k][perm].copy()) for k in ("rY", "rYsrc", "T", "rE", "rEsrc", "mask")}
    got2, _, _ = _gpu_react(capi, mech, "ReactorCvode", ikw, sub)
    assert np.array_equal(got2[0], got[0][:, perm]) and np.array_equal(got2[1], got[1][perm])
    assert np.array_equal(got2[3], got[3][perm])


def test_full_size_properties(capi):
    """BASELINE config 3 size (2^22 cells drm19) would take minutes at tight tolerance inside the unit suite, so the
    full-size pass runs a 2^20 shard and checks the size-independent invariants: rhoH = in + dt*src exactly, total
    mass changes only by dt * sum(src) (elements are conserved by the mechanism), FC > 0 on every unmasked cell,
    FC = 0 and untouched state on masked cells, per-box cost reduction = numpy sum."""
    import torch
    from pelelm_b200 import synth
    mech, nc = "drm19", 1 << 20
    b = synth.synthetic_batch(mech, nc, seed=1, forcing=True, mask_every=97)
    R = capi.Reactor(mech, "ReactorCvode").init(rtol=1e-6, atol=1e-10, max_order=4)
    a = [b["rY"].copy(), b["rYsrc"], b["T"].copy(), b["rE"].copy(), b["rEsrc"], np.zeros(nc)]
    st = R.react([0, 0, 0], [nc - 1, 0, 0], a[0], a[1], a[2], a[3], a[4], a[5], b["mask"], DT)
    m = b["mask"] == -1
    assert st["nfailed"] == 0 and st["nskipped"] == int(m.sum())
    assert np.array_equal(a[3][~m], (b["rE"] + DT * b["rEsrc"])[~m])
    assert np.array_equal(a[0][:, m], b["rY"][:, m]) and np.all(a[5][m] == 0) and np.all(a[5][~m] > 0)
    rho0, rho1 = b["rY"].sum(0), a[0].sum(0)
    drho = (rho1 - rho0 - DT * b["rYsrc"].sum(0))[~m]
    assert np.abs(drho / rho0[~m]).max() < 1e-7
    assert int(a[5].sum()) == st["nfe"]
    # per-box costs on the device
    nb = 64
    off = torch.arange(0, nc + 1, nc // nb, dtype=torch.int64, device="cuda")
    costs = torch.zeros(nb, dtype=torch.int64, device="cuda")
    FCd = torch.from_numpy(a[5]).cuda()
    rc = capi.load().plm_box_costs(FCd.data_ptr(), off.data_ptr(), nb, costs.data_ptr(), None)
    torch.cuda.synchronize()
    assert rc == 0
    assert np.array_equal(costs.cpu().numpy(), a[5].reshape(nb, -1).sum(1).astype(np.int64))
    R.close()


@pytest.mark.parametrize("mech", MECHS)
def test_advance_chemistry_fused_vs_reference(capi, port_lib, mech):
    """plm_advance_chemistry_box = the body of PeleLM::advance_chemistry (PeleLM.cpp:5736-5800) in one kernel pass, against
    the same sequence run around the reference's react (golden) and around the port's react (live): MKS state, I_R,
    bit-exact rhoH, the x1e-3 x1e3 round trip of masked cells, forcing arrays untouched."""
    g = golden_plm(mech)
    ins = {k: g["advchem/in_" + k].copy() for k in ("rhoY", "rhoH", "T", "frcY", "frcH", "mask")}
    nc = ins["T"].size
    R = capi.Reactor(mech, "ReactorCvode").init(rtol=1e-8, atol=1e-10, max_order=2)
    rhoY, rhoH, T = ins["rhoY"].copy(), ins["rhoH"].copy(), ins["T"].copy()
    frcY, frcH = ins["frcY"].copy(), ins["frcH"].copy()
    IR, FC = np.full_like(rhoY, 7.0), np.zeros(nc)
    st = R.advance_chemistry([0, 0, 0], [nc - 1, 0, 0], rhoY, rhoH, T, frcY, frcH, IR, FC, ins["mask"], DT)
    assert not R.cell_status(nc).any()
    R.close()
    assert np.array_equal(frcY, ins["frcY"]) and np.array_equal(frcH, ins["frcH"])
    small = mech != "dodecane_lu"
    ref = (g["advchem/rhoY"], g["advchem/T"], g["advchem/rhoH"], g["advchem/FC"])
    check_state((rhoY, T, rhoH, FC), ref, 1e-7, tag=mech + "/advchem", rtol=1e-6, fc_equal_min=0.9 if small else None,
                fc_rel=0.05 if small else 0.15)
    m = ins["mask"] == -1
    assert m.any() and st["nskipped"] == int(m.sum())
    # masked cells: bit-identical to the reference sequence (scaled to CGS and back, T untouched, I_R from that round trip)
    assert np.array_equal(rhoY[:, m], g["advchem/rhoY"][:, m]) and np.array_equal(T[m], ins["T"][m])
    assert np.array_equal(IR[:, m], g["advchem/IR"][:, m])
    # I_R is exactly -(old - new) * (1/dt) - F of the returned state
    assert np.array_equal(IR, -(ins["rhoY"] - rhoY) * (1.0 / DT) - ins["frcY"])
    e = np.abs(IR - g["advchem/IR"]) / (1e-6 * np.abs(g["advchem/IR"]) + 1e-6 * np.abs(g["advchem/IR"]).max(0, keepdims=True) + 1e-7 / DT)
    assert e.max() <= 1.0, e.max()
    # the same through device pointers, without I_R
    import torch
    R = capi.Reactor(mech, "ReactorCvode").init(rtol=1e-8, atol=1e-10, max_order=2)
    d = {k: torch.from_numpy(v.copy()).cuda() for k, v in ins.items()}
    FCd = torch.zeros(nc, dtype=torch.float64, device="cuda")
    R.advance_chemistry([0, 0, 0], [nc - 1, 0, 0], d["rhoY"], d["rhoH"], d["T"], d["frcY"], d["frcH"], None, FCd, d["mask"], DT)
    torch.cuda.synchronize()
    R.close()
    assert np.array_equal(d["rhoY"].cpu().numpy(), rhoY) and np.array_equal(d["rhoH"].cpu().numpy(), rhoH)
    assert np.array_equal(d["T"].cpu().numpy(), T) and np.array_equal(FCd.cpu().numpy(), FC)


@pytest.mark.parametrize("mech", MECHS)
def test_reaction_rate_rhoY_vs_reference_and_port(capi, port_lib, mech):
    """plm_reaction_rate_rhoY vs the reference's own reactionRateRhoY (Source/PeleLM_K.H:57-101) and the port."""
    g = golden_plm(mech)
    rhoY, rhoH, T, mask = g["rrate/rhoY"], g["rrate/rhoH"], g["rrate/T"], g["rrate/mask"]
    nc = T.size
    R = capi.Reactor(mech, "ReactorCvode").init()
    out = np.full_like(rhoY, -5.0)
    Tin = T.copy()
    R.reaction_rate_rhoY([0, 0, 0], [nc - 1, 0, 0], rhoY.copy(), rhoH.copy(), Tin, mask.copy(), out)
    R.close()
    assert np.array_equal(Tin, T)
    m = mask < 0
    assert m.any() and np.all(out[:, m] == 0.0)
    ref = g["rrate/rhoYdot"]
    err = np.abs(out - ref) / np.abs(ref).max(0, keepdims=True).clip(1e-300)
    assert err[:, ~m].max() < 1e-9, err[:, ~m].max()
    pr = port_lib.Port(mech).reaction_rate_rhoY(rhoY, rhoH, T, mask)
    err = np.abs(out - pr) / np.abs(pr).max(0, keepdims=True).clip(1e-300)
    assert err[:, ~m].max() < 1e-11, err[:, ~m].max()


def test_rk64_two_cells_per_warp_is_bit_identical(capi):
    """PLM_RK64_DUAL=1 selects the two-cells-per-warp ReactorRK64 kernel (rhs_warp2: both right-hand sides at once):
    every cell must come out with exactly the bits of the default one-cell-per-warp kernel, whoever it was paired with."""
    import os
    import subprocess
    import sys
    import tempfile
    from conftest import ROOT
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from pelelm_b200 import capi, synth\n"
        "out = {}\n"
        "for mech, nc in (('LiDryer', 301), ('drm19', 130)):\n"
        "    b = synth.synthetic_batch(mech, nc, seed=17, forcing=True)\n"
        "    R = capi.Reactor(mech, 'ReactorRK64').init(rtol=1e-10, atol=1e-10)\n"
        "    a = [b['rY'].copy(), b['rYsrc'].copy(), b['T'].copy(), b['rE'].copy(), b['rEsrc'].copy(), np.zeros(nc)]\n"
        "    st = R.react([0, 0, 0], [nc - 1, 0, 0], a[0], a[1], a[2], a[3], a[4], a[5], None, 1e-5)\n"
        "    R.close()\n"
        "    out[mech + '_rY'], out[mech + '_T'], out[mech + '_rE'], out[mech + '_FC'] = a[0], a[2], a[3], a[5]\n"
        "    out[mech + '_nfe'] = np.array(st['nfe'])\n"
        "np.savez(sys.argv[1], **out)\n") % ROOT
    res = []
    with tempfile.TemporaryDirectory() as d:
        for dual in ("0", "1"):
            path = os.path.join(d, "o%s.npz" % dual)
            env = dict(os.environ, PLM_RK64_DUAL=dual)
            subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=600)
            res.append(dict(np.load(path)))
    for k in res[0]:
        assert np.array_equal(res[0][k], res[1][k], equal_nan=True), k
    assert res[0]["LiDryer_FC"].min() > 0


@pytest.mark.parametrize("mech,integ", [("drm19", "ReactorCvode"), ("LiDryer", "ReactorRK64")])
def test_subcycled_calls_with_advancing_time(capi, port_lib, mech, integ):
    """ReactEval's loop (PelePhysics/Testing/Exec/ReactEval/main.cpp:520-560, `ode.ndt` sub-steps): react() called
    repeatedly on its own output with time = 0, dt, 2 dt (the forcing term of fKernelSpec uses t - time_init)."""
    from pelelm_b200 import synth
    nc, ndt = 96, 3
    b = synth.synthetic_batch(mech, nc, seed=77, forcing=True, t_range=(900.0, 1400.0) if integ == "ReactorRK64" else (900.0, 2600.0))
    ikw = dict(rtol=1e-9, atol=1e-10, max_order=2)
    P = port_lib.Port(mech).init(integrator=integ, **ikw)
    R = capi.Reactor(mech, integ).init(**ikw)
    g = [b["rY"].copy(), b["rYsrc"].copy(), b["T"].copy(), b["rE"].copy(), b["rEsrc"].copy(), np.zeros(nc)]
    p = [b["rY"].copy(), b["T"].copy(), b["rE"].copy()]
    mask = np.ones(nc, dtype=np.int32)
    for k in range(ndt):
        t = k * DT
        R.react([0, 0, 0], [nc - 1, 0, 0], g[0], g[1], g[2], g[3], g[4], g[5], mask, DT, time=t)
        rY, T, rE, FC = P.react_soa(p[0], b["rYsrc"], p[1], p[2], b["rEsrc"], mask, DT, time=t, nthreads=4)
        p = [rY, T, rE]
        check_state((g[0], g[2], g[3], g[5]), (rY, T, rE, FC), 1e-10, tag="%s step %d" % (mech, k), rtol=2e-6, fc_rel=0.05)
    R.close()
    assert np.array_equal(g[3], p[2])  # rhoH: in + dt*src three times, bit-exact


@pytest.mark.parametrize("mech", MECHS)
def test_internal_energy_reactor_vs_reference(capi, mech):
    """init(reactor_type = 1, ...) (ReactorBase.H:20; the ReactEval default): RHS, Jacobian (consP = 0) and both
    integrators against the reference run with ode.reactor_type = 1."""
    g = golden_e(mech)
    R = capi.Reactor(mech, "ReactorCvode").init(reactor_type=1)
    f = R.eval_rhs(g["rhs/y"], g["rhs/rhoe"], g["rhs/rhoesrc"], g["rhs/rYsrc"], float(g["rhs/t"]))
    assert (np.abs(f - g["rhs/ydot"]) / np.abs(g["rhs/ydot"]).max(1, keepdims=True)).max() < 1e-9
    J = R.eval_jac(g["jac/y"])
    assert (np.abs(J - g["jac/J"]) / np.abs(g["jac/J"]).max(1, keepdims=True)).max() < 1e-9
    R.close()
    small = mech != "dodecane_lu"
    for case, integ, ikw in (("cv_e", "ReactorCvode", dict(rtol=1e-8, atol=1e-10, max_order=2)),
                             ("rk_e", "ReactorRK64", dict(rtol=1e-10, atol=1e-10))):
        if case + "/rY" not in g:
            continue
        b = {k: g[case + "/in_" + k] for k in ("rY", "rYsrc", "T", "rE", "rEsrc", "mask")}
        nc = b["T"].size
        R = capi.Reactor(mech, integ).init(reactor_type=1, **ikw)
        a = [b["rY"].copy(), b["rYsrc"].copy(), b["T"].copy(), b["rE"].copy(), b["rEsrc"].copy(), np.zeros(nc)]
        R.react([0, 0, 0], [nc - 1, 0, 0], a[0], a[1], a[2], a[3], a[4], a[5], b["mask"].copy(), DT)
        assert not R.cell_status(nc).any()
        R.close()
        ref = (g[case + "/rY"], g[case + "/T"], g[case + "/rE"], g[case + "/FC"])
        check_state((a[0], a[2], a[3], a[5]), ref, 1e-10, tag="%s/%s" % (mech, case), rtol=1e-6,
                    fc_equal_min=0.9 if small else None, fc_rel=0.05 if small else 0.2)
    # the enthalpy-only entry point refuses an internal-energy reactor
    R = capi.Reactor(mech, "ReactorCvode").init(reactor_type=1)
    with pytest.raises(Exception):
        z = np.zeros((R.ns, 1))
        R.reaction_rate_rhoY([0, 0, 0], [0, 0, 0], z + 1.0, np.ones(1), np.full(1, 1000.0), None, z.copy())
    R.close()


def test_amr_levels_vs_port(capi, port_lib):
    """BASELINE.json configs[4]: per-level chemistry of the synthetic 3-level TripleFlame hierarchy through
    pelelm_b200/amr.py (costs by the library's box-cost kernel, knapsack, one fused advance_chemistry launch per box on
    several streams, average-down), against the same driver run around the port's react on the CPU for levels 0 and 1
    (they do not depend on level 2's results within three steps except through the average-down, checked separately)."""
    import torch
    from pelelm_b200 import amr
    import test_amr_cpu as tc
    mech, n0, box, nsteps = tc.MECH, (8, 16, 8), 8, 3
    R = capi.Reactor(mech, "ReactorCvode").init(rtol=1e-8, atol=1e-10, max_order=2)
    ns = R.ns
    streams = [torch.cuda.Stream() for _ in range(4)]
    adv, wait = amr.gpu_advance(R, streams)
    dev = amr.build_tripleflame(mech, 1, 0, n0=n0, nlev=3, box=box, device="cuda", forcing=True, dt=DT)
    fc_sums = {}
    fine_first = None
    for step in range(nsteps):
        for L in dev:
            fine = dev[L.lev + 1] if L.lev + 1 < len(dev) else None
            info = amr.advance_chemistry_level(L, DT / 2 ** L.lev, adv, fine=fine, cost_fn=capi.box_costs, wait=wait)
            # the weights of getFuncCountDM: FC = 1 per cell before the first pass, then what the previous pass wrote
            want = [box ** 3] * len(L.boxes) if step == 0 else fc_sums[L.lev]
            assert info["costs"].tolist() == want
            assert info["moved"] == 0 and (info["chem_owner"] == 0).all()
            fc_sums[L.lev] = [int(L.data[b]["FC"].sum().item()) for b in sorted(L.data)]
        if step == 0:
            fine_first = {b: bx["I_R"].clone() for b, bx in dev[2].data.items()}
    torch.cuda.synchronize()
    # one launch per box (warp-per-cell kernels) or one per level (lane = cell kernels: plm_react_boxes)
    assert adv.launches[0] in (nsteps * sum(len(L.boxes) for L in dev), nsteps * len(dev))
    assert not R.cell_status(box ** 3).any()
    R.close()
    # level 1's covered cells after its 3rd advance = mean of level 2's OLD I_R at that moment = level 2's 1st result
    L1, L2 = dev[1], dev[2]
    ncov = 0
    for f, (lo, hi) in enumerate(L2.boxes):
        c = [i for i, (clo, chi) in enumerate(L1.boxes) if all(clo[d] <= lo[d] // 2 and hi[d] // 2 <= chi[d] for d in range(3))]
        assert len(c) == 1
        c = c[0]
        v = fine_first[f].reshape(ns, box // 2, 2, box // 2, 2, box // 2, 2).mean(dim=(2, 4, 6))
        o = [lo[d] // 2 - L1.boxes[c][0][d] for d in range(3)]
        got = L1.data[c]["I_R"].reshape(ns, box, box, box)[:, o[2]:o[2] + box // 2, o[1]:o[1] + box // 2, o[0]:o[0] + box // 2]
        assert torch.equal(got, v)
        ncov += v[0].numel()
    assert 0 < ncov < L1.ncells()
    # levels 0 and 1 against the port: uncovered cells only (covered I_R comes from the finer level, checked above)
    cpu = amr.build_tripleflame(mech, 1, 0, n0=n0, nlev=3, box=box, forcing=True, dt=DT)
    padv = tc._port_advance()
    for _ in range(nsteps):
        for L in cpu[:2]:
            amr.advance_chemistry_level(L, DT / 2 ** L.lev, padv, fine=None, use_knapsack=False)
    for l in range(2):
        fineb = dev[l + 1].boxes
        for b in cpu[l].data:
            g = {k: dev[l].data[b][k].cpu().numpy() for k in ("S", "I_R", "FC")}
            r = {k: cpu[l].data[b][k].numpy() for k in ("S", "I_R", "FC")}
            ref = (r["S"][:ns], r["S"][ns + 1], r["S"][ns], r["FC"])
            got = (g["S"][:ns], g["S"][ns + 1], g["S"][ns], g["FC"])
            check_state(got, ref, 1e-7, tag="amr L%d box %d" % (l, b), rtol=1e-6, fc_rel=0.05)
            cov = np.zeros((box, box, box), dtype=bool)
            blo = cpu[l].boxes[b][0]
            for lo, hi in fineb:
                o = [lo[d] // 2 - blo[d] for d in range(3)]
                if all(0 <= o[d] < box for d in range(3)):
                    cov[o[2]:o[2] + box // 2, o[1]:o[1] + box // 2, o[0]:o[0] + box // 2] = True
            unc = ~cov.ravel()
            sc = 1e-6 * np.abs(r["I_R"]) + 1e-6 * np.abs(r["I_R"]).max(1, keepdims=True) + 1e-7 / DT
            assert (np.abs(g["I_R"] - r["I_R"]) <= sc)[:, unc].all(), (l, b)


@pytest.mark.parametrize("mech", ["drm19", "dodecane_lu"])
def test_flamesheet_profile_vs_reference(capi, mech):
    """BASELINE.json configs[0]: states across the FlameSheet case's premixed-flame profile with the chemistry settings of
    Exec/RegTests/FlameSheet/inputs.2d-regt (rtol 1e-6, atol 1e-5 x typical values, order 4), CUDA path against the
    reference's own output (tests/golden/<mech>_flamesheet.npz, make_golden.py::flamesheet)."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, mech + "_flamesheet.npz"))
    ikw = dict(rtol=1e-6, atol=1e-5, max_order=4)
    b = {k: g["in_" + k] for k in ("rY", "rYsrc", "T", "rE", "rEsrc", "mask")}
    got, st, status = _gpu_react(capi, mech, "ReactorCvode", ikw, b, g["typ"])
    ref = (g["rY"], g["T"], g["rE"], g["FC"])
    small = mech != "dodecane_lu"
    check_state(got, ref, ikw["atol"] * g["typ"][:-1, None], tag=mech + "/flamesheet", rtol=100 * ikw["rtol"],
                fc_equal_min=0.9 if small else None, fc_rel=0.05 if small else 0.15)
    assert not status.any() and st["nskipped"] == 0


@pytest.mark.parametrize("mech,path", [("LiDryer", "auto"), ("drm19", "auto"), ("drm19", "lane")])
def test_react_boxes_equals_per_box_calls(capi, mech, path):
    """plm_react_boxes (every local box of a level in one call; one launch with the lane = cell kernels) gives bit for bit what
    one plm_advance_chemistry_box call per box gives: a cell's result depends on its own inputs only.  Boxes of different sizes,
    one with ghost cells around it, masks, I_R."""
    import torch
    from pelelm_b200 import synth
    ns = synth.mech(mech).ns
    R = capi.Reactor(mech, "ReactorCvode").init(rtol=1e-8, atol=1e-10, max_order=2).set_path(path)
    shapes = [((0, 0, 0), (15, 7, 3), None, None), ((4, 2, 1), (35, 9, 2), (0, 0, 0), (39, 11, 3)), ((0, 0, 0), (96, 0, 0), None, None)]
    boxes, ref = [], []
    for i, (lo, hi, flo, fhi) in enumerate(shapes):
        flo, fhi = flo or lo, fhi or hi
        nf = int(np.prod([fhi[d] - flo[d] + 1 for d in range(3)]))
        b = synth.synthetic_batch(mech, nf, seed=100 + i, forcing=True, mask_every=13)
        mk = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
        arrs = {"rY": mk(b["rY"] * 1e3), "rYsrc": mk(b["rYsrc"] * 1e3), "T": mk(b["T"]), "rEner": mk(b["rE"] * 0.1),
                "rEner_src": mk(b["rEsrc"] * 0.1), "FC": torch.full((nf,), -3.0, dtype=torch.float64, device="cuda"),
                "I_R": torch.zeros((ns, nf), dtype=torch.float64, device="cuda"), "mask": mk(b["mask"])}
        one = {k: v.clone() for k, v in arrs.items()}
        R.advance_chemistry(list(lo), list(hi), one["rY"], one["rEner"], one["T"], one["rYsrc"], one["rEner_src"], one["I_R"], one["FC"],
                            one["mask"], DT, fab_lo=list(flo), fab_hi=list(fhi))
        ref.append(one)
        boxes.append(dict(arrs, lo=lo, hi=hi, fab_lo=flo, fab_hi=fhi))
    st = R.react_boxes(boxes, DT, mks=True)
    assert st["ncells"] + st["nskipped"] == sum(int(np.prod([hi[d] - lo[d] + 1 for d in range(3)])) for lo, hi, _, _ in shapes)
    assert st["nfailed"] == 0
    for bx, one in zip(boxes, ref):
        for k in ("rY", "T", "rEner", "FC", "I_R"):
            assert torch.equal(bx[k], one[k]), k
    assert (boxes[1]["FC"] == -3.0).any() and (boxes[1]["FC"] > 0).any()   # ghost cells untouched, interior advanced
    R.close()
