"""The lane = cell integrators and the generated straight-line mechanism code (pelelm_b200/csrc/lane_*.h, gen/lane_<mech>.h),
compiled for the HOST as a one-lane model (tests/lane_model.cpp -- test infrastructure, never part of libplmchem.so) and
checked against the UNMODIFIED reference through the golden fixtures and against the oracle port.  The product compiles
the same headers with nvcc for sm_100a; the `-m gpu` suite holds that build to the same fixtures through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, GOLDEN, golden, check_state, rel_errors
from golden.make_golden import CASES, DT

MECHS2 = ["LiDryer", "drm19"]   # the mechanisms the lane = cell kernels cover
# both forms of the mechanism code: "" = straight-line generated code (lane_chem.h), "_tab" = loops over constant tables (lane_tab.h)
FORMS = ["", "_tab"]
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(dp)


@pytest.fixture(scope="session")
def lane_lib():
    d = os.path.join(ROOT, "tests", "lane")
    os.makedirs(d, exist_ok=True)
    so = os.path.join(d, "liblane_model.so")
    csrc = os.path.join(ROOT, "pelelm_b200", "csrc")
    deps = [os.path.join(ROOT, "tests", "lane_model.cpp")] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.startswith("lane_")]
    deps += [os.path.join(csrc, "gen", "lane_%s.h" % m) for m in MECHS2] + [os.path.join(csrc, "gen", "lanetab_%s.h" % m) for m in MECHS2]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-I" + csrc, deps[0], "-o", so])
    L = C.CDLL(so)
    L.lane_model_rhs.argtypes = [C.c_char_p, C.c_int, C.c_double, dp, dp, C.c_double, C.c_double, dp]
    L.lane_model_jac.argtypes = [C.c_char_p, dp, dp]
    L.lane_model_react.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, dp, C.c_int, C.c_int,
                                   C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, ip, C.c_double, C.c_double, C.c_int, C.c_int,
                                   C.c_int, dp, ip]
    return L


def lane_react(L, mech, integ, ikw, b, typ=None, rt=2, time=0.0):
    nc = b["T"].size
    rY, T, rE, FC = b["rY"].copy(), b["T"].copy(), b["rE"].copy(), np.zeros(nc)
    st = np.zeros(8, dtype=np.int32)
    mask = np.ascontiguousarray(b["mask"], dtype=np.int32)
    tv = None if typ is None else _p(np.ascontiguousarray(typ, dtype=np.float64))
    L.lane_model_react(mech.encode(), 1 if integ == "ReactorRK64" else 0, rt, ikw.get("rtol", ikw["atol"]), ikw["atol"],
                       ikw.get("max_order", 2), 3, tv, 10, 5, 500, nc, _p(rY), _p(np.ascontiguousarray(b["rYsrc"])), _p(T), _p(rE),
                       _p(np.ascontiguousarray(b["rEsrc"])), _p(FC), mask.ctypes.data_as(ip), DT, time, 0, 0,
                       ikw.get("clean_init_massfrac", 0), None, st.ctypes.data_as(ip))
    return (rY, T, rE, FC), st


@pytest.mark.parametrize("form", FORMS)
@pytest.mark.parametrize("mech", MECHS2)
@pytest.mark.parametrize("case", [c for c in CASES if c != "cv_dq"] + ["reval"])
def test_lane_react_vs_reference(lane_lib, mech, case, form):
    g = golden(mech)
    integ, ikw, _ = CASES.get(case, ("ReactorCvode", dict(rtol=1e-10, atol=1e-10, max_order=2), None))
    b = {k: g[case + "/in_" + k] for k in ("rY", "rYsrc", "T", "rE", "rEsrc", "mask")}
    typ = g[case + "/typ"] if case + "/typ" in g else None
    got, st = lane_react(lane_lib, mech + form, integ, ikw, b, typ)
    atol = ikw["atol"] * (typ[:-1, None] if typ is not None else 1.0)
    ref = (g[case + "/rY"], g[case + "/T"], g[case + "/rE"], g[case + "/FC"])
    check_state(got, ref, atol, tag="%s/%s" % (mech, case), rtol=max(1e-6, 100 * ikw.get("rtol", 0.0)), fc_equal_min=0.9, fc_rel=0.05)
    assert st[7] == 0


@pytest.mark.parametrize("form", FORMS)
@pytest.mark.parametrize("mech", MECHS2)
def test_lane_counters_equal_port(lane_lib, port_lib, mech, form):
    """the state machine takes exactly the steps, Newton iterations, setups and Jacobians the per-cell CVODE restatement takes"""
    from pelelm_b200 import synth
    b = synth.synthetic_batch(mech, 24, seed=5, forcing=True, mask_every=7)
    ikw = dict(rtol=1e-10, atol=1e-10, max_order=2)
    P = port_lib.Port(mech).init(**ikw)
    ref = P.react_soa(b["rY"], b["rYsrc"], b["T"], b["rE"], b["rEsrc"], b["mask"], DT, return_stats=True)
    got, st = lane_react(lane_lib, mech + form, "ReactorCvode", ikw, b)
    assert np.array_equal(got[3], ref[3])
    for i, k in enumerate(("nst", "nfe", "nje", "nsetups", "nni", "ncfn", "netf")):
        assert st[i] == ref[4][k], (k, st[i], ref[4][k])
    check_state(got, ref[:4], 1e-10, tag=mech)


@pytest.mark.parametrize("form", FORMS)
@pytest.mark.parametrize("mech", MECHS2)
def test_lane_big_batch_vs_reference(lane_lib, mech, form):
    """the headline configuration on the first 256 cells of tests/golden/<mech>_big.npz"""
    from pelelm_b200 import synth
    g = np.load(os.path.join(GOLDEN, mech + "_big.npz"))
    n, ch = 256, int(g["chunk"])
    parts = [synth.synthetic_batch(mech, ch, cell0=c, seed=synth.DEFAULT_SEED + 301, dt=DT, forcing=True, mask_every=61)
             for c in range(0, n, ch)]
    b = {k: np.ascontiguousarray(np.concatenate([p[k] for p in parts], axis=-1)) for k in parts[0]}
    assert np.array_equal(b["T"], g["in_T"][:n])
    got, st = lane_react(lane_lib, mech + form, "ReactorCvode", dict(rtol=1e-10, atol=1e-10, max_order=2), b)
    ref = (np.ascontiguousarray(g["rY"][:, :n]), g["T"][:n], g["rE"][:n], g["FC"][:n])
    check_state(got, ref, 1e-10, tag=mech + "/big", fc_equal_min=0.98, fc_rel=0.01)
    r1, r2 = rel_errors(got[0], ref[0])
    print("%s: max relative error of rhoY_k over Y_k > 1e-12: %.2e; over rhoY_k > atol: %.2e" % (mech, r1, r2))
    assert r2 <= 1e-4


@pytest.mark.parametrize("form", FORMS)
@pytest.mark.parametrize("mech", MECHS2 + ["dodecane_lu"])
def test_lane_rhs_jac_vs_reference(lane_lib, mech, form):
    if mech == "dodecane_lu" and form == "":
        pytest.skip("the straight-line form of dodecane_lu is not compiled into the host model")
    g = golden(mech)
    y, J = g["rhs/y"], g["jac/y"]
    f = np.zeros_like(y)
    for c in range(y.shape[0]):
        lane_lib.lane_model_rhs((mech + form).encode(), 2, float(g["rhs/t"]), _p(np.ascontiguousarray(y[c])), _p(f[c]), g["rhs/rhoe"][c],
                                g["rhs/rhoesrc"][c], _p(np.ascontiguousarray(g["rhs/rYsrc"][c])))
    assert (np.abs(f - g["rhs/ydot"]) / np.abs(g["rhs/ydot"]).max(1, keepdims=True)).max() < 1e-9
    Jg = np.zeros_like(g["jac/J"])
    for c in range(J.shape[0]):
        lane_lib.lane_model_jac((mech + form).encode(), _p(np.ascontiguousarray(J[c])), _p(Jg[c]))
    assert (np.abs(Jg - g["jac/J"]) / np.abs(g["jac/J"]).max(1, keepdims=True)).max() < 1e-11


@pytest.mark.parametrize("form", FORMS)
@pytest.mark.parametrize("mech", MECHS2)
def test_lane_flat_api_vs_reference(lane_lib, mech, form):
    """1dArray API: one cell per call is the reference's per-cell problem"""
    g = np.load(os.path.join(GOLDEN, mech + "_flat.npz"))
    for name, integ in (("cv", 0), ("rk", 1)):
        y, src = g[name + "/in_y"].copy(), np.ascontiguousarray(g[name + "/in_src"])
        rX, rXs = g[name + "/in_rX"].copy(), np.ascontiguousarray(g[name + "/in_rXsrc"])
        nc = rX.size
        FC = np.zeros(nc)
        st = np.zeros(8, dtype=np.int32)
        lane_lib.lane_model_react((mech + form).encode(), integ, 2, 1e-10, 1e-10, 2, 3, None, 10, 5, 500, nc, _p(y), _p(src), None, _p(rX),
                                  _p(rXs), _p(FC), None, DT, 0.0, 1, 0, 0, None, st.ctypes.data_as(ip))
        ok = ~np.isnan(g[name + "/y"]).any(1)
        assert np.array_equal(ok, ~np.isnan(y).any(1))
        assert np.array_equal(rX[ok], g[name + "/rX"][ok])
        if not ok.any():   # the explicit integrator blows up on every stiff state of this mechanism: same NaN cells, done
            continue
        e = np.abs(y - g[name + "/y"])[ok] / (1e-6 * np.abs(g[name + "/y"][ok]) + 1e-10)
        assert e.max() <= 1.0, (name, e.max())
        assert np.mean(FC[ok] == g[name + "/FC"][ok]) >= 0.9, (name, FC, g[name + "/FC"])
