"""Build libplmchem.so (sm_100a only) in-tree with nvcc. Used by __graft_entry__.build() and the tests."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# PLM_VARIANT (development only): build a differently-flagged copy next to the product library, e.g.
#   PLM_VARIANT=mb6 PLM_EXTRA_CUFLAGS=-DPLM_MIN_BLOCKS=6 python -m pelelm_b200.build ; PLM_LIBPATH=... python bench.py
VARIANT = os.environ.get("PLM_VARIANT", "")
LIB = os.path.join(HERE, "libplmchem%s.so" % (("_" + VARIANT) if VARIANT else ""))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CUFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + os.environ.get("PLM_EXTRA_CUFLAGS", "").split()

SOURCES = ["kernels.cu", "capi.cpp", "dev_tables.cpp", "knapsack.cpp", "lane_dispatch.cpp", "gen/tables_LiDryer.c",
           "gen/tables_drm19.c", "gen/tables_dodecane_lu.c"]
# the lane = cell kernels: one translation unit per (mechanism, kernel family) so the straight-line code compiles in parallel
LANE_UNITS = [(m, k) for m in ("LiDryer", "drm19") for k in (0, 2, 5)]
# form of the mechanism code per mechanism: 1 = loops over constant-memory tables (lane_tab.h), 0 = straight-line generated
# code (lane_chem.h).  Measured on B200 (profiles/r02c_lane_ab*.json): the loop form executes ~3x the instructions of the
# straight-line form (index loads, address arithmetic, uniform branches) and is 3x slower for LiDryer, 1.7x slower for drm19.
# PLM_LANE_TAB="LiDryer=1,drm19=0" overrides (development A/B, with PLM_VARIANT).
LANE_TAB = {"LiDryer": 0, "drm19": 0}
for _kv in filter(None, os.environ.get("PLM_LANE_TAB", "").split(",")):
    LANE_TAB[_kv.split("=")[0]] = int(_kv.split("=")[1])


def _newer(src: str, dst: str) -> bool:
    return (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst)


def _compile(cmd, log=None):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("compile failed: " + " ".join(cmd))
    if log:
        with open(log, "w") as f:
            f.write(r.stdout + r.stderr)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(CSRC, "build" + (("_" + VARIANT) if VARIANT else ""))
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers += [os.path.join(CSRC, "gen", "mech_tables.h"), os.path.join(os.path.dirname(HERE), "include", "plm_chem.h")]
    # the warp-per-cell kernels do not include the lane headers and vice versa: separate time stamps keep rebuilds short
    lane_hdr = [h for h in headers if os.path.basename(h).startswith("lane_")]
    hdr_time = max(os.path.getmtime(h) for h in headers if h not in lane_hdr)
    jobs = []
    objs = []
    relink = force or not os.path.exists(LIB)
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(obj)
        if force or _newer(src, obj) or os.path.getmtime(obj) < hdr_time:
            if s.endswith(".c"):
                cmd = ["gcc", "-O2", "-fPIC", "-I" + CSRC, "-I" + os.path.join(CSRC, "gen"), "-c", src, "-o", obj]
            elif s.endswith(".cu"):
                cmd = [NVCC] + ARCH + CUFLAGS + ["-I" + CSRC, "-x", "cu", "-c", src, "-o", obj]
            else:
                cmd = [NVCC] + ARCH + ["-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-I" + CSRC, "-c", src, "-o", obj]
            log = os.path.join(objdir, os.path.basename(s) + ".ptxas.log") if (verbose or s.endswith(".cu")) else None
            jobs.append((cmd, log))
    lane_src = os.path.join(CSRC, "lane_inst.cu")
    for mech, kind in LANE_UNITS:
        gen = os.path.join(CSRC, "gen", ("lanetab_%s.h" if LANE_TAB[mech] else "lane_%s.h") % mech)
        obj = os.path.join(objdir, "lane_%s_%d.o" % (mech, kind))
        objs.append(obj)
        t = max([os.path.getmtime(h) for h in lane_hdr] + [os.path.getmtime(gen), os.path.getmtime(lane_src),
                                                           os.path.getmtime(os.path.join(CSRC, "chem_warp.cuh")),
                                                           os.path.getmtime(os.path.join(CSRC, "kernels.h"))])
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < t:
            cmd = [NVCC] + ARCH + CUFLAGS + ["-I" + CSRC, "-DLANE_MECH=" + mech, "-DLANE_KIND=%d" % kind,
                                           "-DLANE_TAB=%d" % LANE_TAB[mech], "-c", lane_src, "-o", obj]
            jobs.append((cmd, os.path.join(objdir, "lane_%s_%d.ptxas.log" % (mech, kind))))
    if jobs:
        with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 4))) as ex:
            list(ex.map(lambda j: _compile(*j), jobs))
        relink = True
    if relink:
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
