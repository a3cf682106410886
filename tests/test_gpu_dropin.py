"""Drop-in check on the GPU box: the UNMODIFIED reference driver (oracle/_ref/libref_<mech>_b200.so = reference
objects + pelelm_b200/adaptor/ReactorB200.cpp) selects `chem_integrator = Reactor{Cvode,RK64}B200` through its own
factory (Factory.H:73-80) and calls the virtual react(box, ...) exactly as PeleLM does (PeleLM.cpp:5752-5758); the
result is compared with the stock CPU reactor run in the same process on the same FAB."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, check_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mech,integ", [("drm19", "ReactorCvode"), ("LiDryer", "ReactorCvode"),
                                        ("LiDryer", "ReactorRK64"), ("dodecane_lu", "ReactorCvode")])
def test_reference_factory_creates_b200_reactor(tmp_path, plm_lib, mech, integ):
    from oracle import pyoracle
    if not pyoracle.ref_available(mech, b200=True):
        pytest.skip("oracle/_ref/libref_%s_b200.so not built (needs /root/reference at build time)" % mech)
    out = str(tmp_path / "r.npz")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_worker.py"), ROOT, mech, integ, out],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    g = np.load(out)
    a, b = integ, integ + "B200"
    small = mech != "dodecane_lu"
    check_state((g[b + "/rY"], g[b + "/T"], g[b + "/rE"], g[b + "/FC"]),
                (g[a + "/rY"], g[a + "/T"], g[a + "/rE"], g[a + "/FC"]), 1e-10, tag="%s/%s" % (mech, integ),
                fc_equal_min=0.9 if small else None, fc_rel=0.05 if small else 0.15)


def _read_plt(path):
    """every level of a PeleLM plotfile: component names and, per level, the list of FABs [ncomp, nz, ny, nx]"""
    from pelelm_b200 import vismf
    hdr = open(os.path.join(path, "Header")).read().split("\n")
    ncomp = int(hdr[1])
    names = hdr[2:2 + ncomp]
    levels = []
    lev = 0
    while os.path.isdir(os.path.join(path, "Level_%d" % lev)):
        levels.append(vismf.read_multifab(os.path.join(path, "Level_%d" % lev, "Cell"))["fabs"])
        lev += 1
    return names, levels


def test_flamesheet_run_through_the_adaptor(tmp_path, plm_lib):
    """BASELINE.json configs[0] end to end: the reference's own Exec/RegTests/FlameSheet executable (built by
    tools/flamesheet/build.sh with ReactorB200.cpp + -lplmchem; CPU AMReX, so PeleLM's overlapping views of one STemp FAB go
    through the host-pointer path), `inputs.2d-regt max_step=3`, once with the stock `ReactorCvode` and once with
    `peleLM.chem_integrator = ReactorCvodeB200` (Exec/RegTests/FlameSheet/inputs.2d-regt:86, Source/PeleLM.cpp:701-709).
    Plotfiles of the two runs are compared; TinyProfiler times of PLM::advance_chemistry are printed."""
    import re
    import shutil
    d = os.path.join(ROOT, "oracle", "_ref", "flamesheet")
    exe = os.path.join(d, "PeleLM2d.gnu.TPROF.ex")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/flamesheet not built (tools/flamesheet/build.sh; needs /root/reference)")
    for f in ("inputs.2d-regt", "drm19_pmf.dat"):
        shutil.copy(os.path.join(d, f), str(tmp_path / f))
    times = {}
    for name, integ in (("ref", "ReactorCvode"), ("b200", "ReactorCvodeB200")):
        r = subprocess.run([exe, "inputs.2d-regt", "max_step=3", "peleLM.chem_integrator=" + integ, "amr.plot_file=plt_" + name,
                            "amr.plot_int=3", "amr.check_int=-1"], cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
        logdir = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(logdir):  # the full run log travels back from the GPU box
            with open(os.path.join(logdir, "flamesheet_%s.log" % name), "w") as f:
                f.write(r.stdout + "\n--- stderr ---\n" + r.stderr + "\n--- files ---\n" + "\n".join(sorted(os.listdir(str(tmp_path)))))
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
        assert os.path.exists(str(tmp_path / ("plt_%s00003" % name) / "Header")), \
            "no plotfile after 3 steps:\n" + r.stdout[:3000] + "\n...\n" + r.stdout[-3000:] + r.stderr[-2000:]
        m = re.findall(r"PLM::advance_chemistry\(\)\s+\d+\s+([0-9.eE+-]+)\s+([0-9.eE+-]+)\s+([0-9.eE+-]+)\s+([0-9.]+)%", r.stdout)
        times[name] = float(m[-1][2]) if m else None   # inclusive table: max time over ranks
    na, la = _read_plt(str(tmp_path / "plt_ref00003"))
    nb, lb = _read_plt(str(tmp_path / "plt_b20000003"))
    assert na == nb and len(la) == len(lb)
    iT, iFC = na.index("temp"), na.index("FuncCount")
    ys = [i for i, n in enumerate(na) if n.startswith("Y(")]
    eT, eY, fa, fb = 0.0, 0.0, 0.0, 0.0
    for A, B in zip(la, lb):
        assert len(A) == len(B)
        for x, y in zip(A, B):
            assert x.shape == y.shape
            eT = max(eT, float(np.max(np.abs(y[iT] / x[iT] - 1.0))))
            eY = max(eY, float(np.max(np.abs(y[ys] - x[ys]))))
            fa += float(x[iFC].sum())
            fb += float(y[iFC].sum())
    print("FlameSheet 3 steps: max |dT|/T %.2e, max |dY_k| %.2e, sum FuncCount %.0f (reference) vs %.0f (B200); "
          "PLM::advance_chemistry %s s (reference, 1 core) vs %s s (B200 through the adaptor)" % (eT, eY, fa, fb, times["ref"], times["b200"]))
    # chemistry runs at rtol 1e-6, atol 1e-5 x typical values; the flow solver couples the cells for three steps
    assert eT <= 1e-4 and eY <= 1e-5
    assert abs(fb / fa - 1.0) <= 0.05
