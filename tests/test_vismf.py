"""pelelm_b200/vismf.py against the reference's own VisMF (fixture written by amrex::VisMF::Write through the reference
driver, tests/golden/make_golden.py::vismf) and the ReactEval chkfile round trip (main.cpp:206-300, :358-405)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from pelelm_b200 import vismf


def test_reader_matches_reference_written_multifab():
    vals = np.load(os.path.join(GOLDEN, "vismf", "values.npy"))
    mf = vismf.read_multifab(os.path.join(GOLDEN, "vismf", "Level_0", "OldState"))
    nb, nc, side = vals.shape[0], vals.shape[1], vals.shape[2]
    assert mf["ncomp"] == nc and mf["ngrow"] == 0 and len(mf["fabs"]) == nb
    for b in range(nb):
        assert mf["boxes"][b] == ((b * side, 0, 0), ((b + 1) * side - 1, side - 1, side - 1))
        assert np.array_equal(mf["fabs"][b], vals[b])  # bit-exact


def test_writer_output_is_byte_identical_to_the_reference(tmp_path):
    """Same MultiFab written by write_multifab: header and data files equal the reference's byte for byte."""
    vals = np.load(os.path.join(GOLDEN, "vismf", "values.npy"))
    ref = vismf.read_multifab(os.path.join(GOLDEN, "vismf", "Level_0", "OldState"))
    vismf.write_multifab(str(tmp_path / "Level_0" / "OldState"), ref["boxes"], [vals[b] for b in range(vals.shape[0])])
    for name in ("OldState_H", "OldState_D_00000"):
        a = open(os.path.join(GOLDEN, "vismf", "Level_0", name), "rb").read()
        b = open(str(tmp_path / "Level_0" / name), "rb").read()
        assert a == b, name


def test_reference_reads_what_we_write(tmp_path):
    """amrex::VisMF::Read (the reference, oracle/_ref) on a file written by write_multifab."""
    from oracle import pyoracle
    if not pyoracle.ref_available("LiDryer"):
        pytest.skip("oracle/_ref not built here")
    rng = np.random.default_rng(3)
    boxes = [((0, 0, 0), (7, 3, 1)), ((8, 0, 0), (11, 3, 1)), ((0, 4, 0), (11, 5, 1))]
    fabs = [rng.normal(size=(4, hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1)) for lo, hi in boxes]
    prefix = str(tmp_path / "Level_0" / "ChemForcing")
    vismf.write_multifab(prefix, boxes, fabs)
    R = pyoracle.Ref("LiDryer")
    R.L.ref_vismf_read.restype = C.c_long
    R.L.ref_vismf_read.argtypes = [C.c_char_p, C.c_long, C.c_void_p, C.c_void_p, C.c_int]
    out = np.zeros(sum(f.size for f in fabs))
    meta = np.zeros(2 + 6 * len(boxes), dtype=np.int32)
    n = R.L.ref_vismf_read(prefix.encode(), out.size, out.ctypes.data, meta.ctypes.data, meta.size)
    assert n == out.size and meta[0] == len(boxes) and meta[1] == 4
    off = 0
    for b, ((lo, hi), f) in enumerate(zip(boxes, fabs)):
        assert tuple(meta[2 + 6 * b:5 + 6 * b]) == lo and tuple(meta[5 + 6 * b:8 + 6 * b]) == hi
        assert np.array_equal(out[off:off + f.size], f.ravel())
        off += f.size


def test_reacteval_chk_round_trip(tmp_path):
    """A synthetic PeleLM-style dump (MKS, velocity + rho + rhoY + rhoh + T + rhoRT) comes back as the CGS arrays
    ReactEval would integrate (main.cpp:388-399), box by box and level by level."""
    from pelelm_b200 import synth
    mech, ns, dim = "drm19", 21, 3
    levels, expect = [], []
    for lev, boxes in enumerate([[((0, 0, 0), (3, 3, 1)), ((4, 0, 0), (7, 3, 1))], [((2, 2, 0), (5, 5, 3))]]):
        L, E = [], []
        for bi, (lo, hi) in enumerate(boxes):
            shp = (hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1)
            nc = int(np.prod(shp))
            b = synth.synthetic_batch(mech, nc, seed=100 * lev + bi, forcing=True)
            state = np.zeros((dim + ns + 4,) + shp)
            state[dim] = (b["rY"].sum(0) * 1e3).reshape(shp)
            state[dim + 1:dim + 1 + ns] = (b["rY"] * 1e3).reshape((ns,) + shp)
            state[dim + ns + 1] = (b["rE"] * 0.1).reshape(shp)
            state[dim + ns + 2] = b["T"].reshape(shp)
            forcing = np.concatenate([(b["rYsrc"] * 1e3).reshape((ns,) + shp), (b["rEsrc"] * 0.1).reshape((1,) + shp)])
            L.append({"lo": lo, "hi": hi, "state": state, "forcing": forcing})
            E.append((state, forcing))
        levels.append(L)
        expect.append(E)
    chk = str(tmp_path / "chk00010")
    vismf.write_reacteval_chk(chk, 2.5e-6, [0.0, 0.0, 0.0], [0.008, 0.004, 0.002], ((0, 0, 0), (7, 3, 1)), levels, ns)
    got = vismf.read_reacteval_chk(chk, ns)
    assert got["dt"] == 2.5e-6 and got["finest_level"] == 1 and got["domain"] == ((0, 0, 0), (7, 3, 1))
    for lev in range(2):
        assert len(got["levels"][lev]) == len(levels[lev])
        for g, src, (state, forcing) in zip(got["levels"][lev], levels[lev], expect[lev]):
            assert g["lo"] == tuple(src["lo"]) and g["hi"] == tuple(src["hi"])
            assert np.array_equal(g["rY"], state[dim + 1:dim + 1 + ns] * 1.0e-3)
            assert np.array_equal(g["rYsrc"], forcing[:ns] * 1.0e-3)
            assert np.array_equal(g["T"], state[dim + ns + 2])
            assert np.array_equal(g["rE"], state[dim + ns + 1] * 10.0)
            assert np.array_equal(g["rEsrc"], forcing[ns] * 10.0)
    with pytest.raises(ValueError):
        vismf.read_reacteval_chk(chk, ns + 1)


@pytest.mark.gpu
def test_replay_of_a_state_dump_through_the_cuda_path(tmp_path, port_lib):
    """ReactEval flow on a dump: load, react box by box on the GPU, compare with the oracle port on the same arrays."""
    from pelelm_b200 import capi, synth
    from conftest import check_state
    mech, ns, dim = "LiDryer", 9, 3
    lo, hi = (0, 0, 0), (7, 3, 1)
    shp = (2, 4, 8)
    nc = 64
    b = synth.synthetic_batch(mech, nc, seed=5, forcing=True)
    state = np.zeros((dim + ns + 4,) + shp)
    state[dim + 1:dim + 1 + ns] = (b["rY"] * 1e3).reshape((ns,) + shp)
    state[dim + ns + 1] = (b["rE"] * 0.1).reshape(shp)
    state[dim + ns + 2] = b["T"].reshape(shp)
    forcing = np.concatenate([(b["rYsrc"] * 1e3).reshape((ns,) + shp), (b["rEsrc"] * 0.1).reshape((1,) + shp)])
    chk = str(tmp_path / "chk")
    vismf.write_reacteval_chk(chk, 1e-5, [0, 0, 0], [1, 1, 1], (lo, hi), [[{"lo": lo, "hi": hi, "state": state, "forcing": forcing}]], ns)
    d = vismf.read_reacteval_chk(chk, ns)
    bx = d["levels"][0][0]
    a = [np.ascontiguousarray(bx[k]) for k in ("rY", "rYsrc", "T", "rE", "rEsrc")]
    flat = [x.reshape(ns, -1).copy() if x.ndim == 4 else x.reshape(-1).copy() for x in a]
    FC = np.zeros(shp)
    R = capi.Reactor(mech, "ReactorCvode").init(rtol=1e-8, atol=1e-10, max_order=2)
    R.react(list(bx["lo"]), list(bx["hi"]), a[0], a[1], a[2], a[3], a[4], FC, None, d["dt"])
    R.close()
    P = port_lib.Port(mech).init(rtol=1e-8, atol=1e-10, max_order=2)
    ref = P.react_soa(flat[0], flat[1], flat[2], flat[3], flat[4], np.ones(nc, dtype=np.int32), d["dt"], nthreads=4)
    check_state((a[0].reshape(ns, -1), a[2].reshape(-1), a[3].reshape(-1), FC.reshape(-1)), ref, 1e-10, tag="replay",
                rtol=1e-6, fc_equal_min=0.9)
