"""CPU-side checks of the product library: it builds, loads, exports the whole C ABI, refuses to compute
without a GPU (no CPU fallback), and its host-only cost model (knapsack) matches the reference."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, GOLDEN


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "plm_chem.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(plm_[A-Za-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported(plm_lib):
    syms = _declared_symbols()
    assert len(syms) >= 14
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(ROOT, "pelelm_b200", "libplmchem.so")],
                                  text=True)
    exported = set(l.split()[-1] for l in out.splitlines() if l.strip())
    missing = [s for s in syms if s not in exported]
    assert not missing, missing
    for s in syms:
        getattr(plm_lib, s)


def test_binding_covers_header(plm_lib):
    from pelelm_b200 import capi
    assert sorted(capi.EXPORTS) == _declared_symbols()


def test_sass_is_sm100a_only():
    out = subprocess.check_output(["/usr/local/cuda/bin/cuobjdump", "-lelf",
                                   os.path.join(ROOT, "pelelm_b200", "libplmchem.so")], text=True)
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_oracle_in_product():
    """The product tree must not reference the oracle (parity claims are void otherwise)."""
    for dp, _, fns in os.walk(os.path.join(ROOT, "pelelm_b200")):
        if "build" in dp.split(os.sep):
            continue
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "libref_" not in txt, os.path.join(dp, fn)
    out = subprocess.check_output(["ldd", os.path.join(ROOT, "pelelm_b200", "libplmchem.so")], text=True)
    assert "oracle" not in out and "libref" not in out


def test_fails_loudly_without_gpu(plm_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pelelm_b200 import capi
    with pytest.raises(capi.PlmError) as ei:
        capi.Reactor("drm19", "ReactorCvode")
    assert "no CUDA device" in str(ei.value) or "CPU path" in str(ei.value)


def test_bad_names_rejected(plm_lib):
    h = C.c_void_p()
    assert plm_lib.plm_reactor_create(C.byref(h), b"nosuchmech", b"ReactorCvode") == -4
    assert b"valid" in plm_lib.plm_last_error()
    rc = plm_lib.plm_reactor_create(C.byref(h), b"drm19", b"ReactorArkode")
    assert rc == -1 and b"ReactorCvode, ReactorRK64" in plm_lib.plm_last_error()


def test_knapsack_vs_reference(plm_lib):
    from pelelm_b200 import capi
    g = np.load(os.path.join(GOLDEN, "knapsack.npz"))
    i = 0
    while "w%d" % i in g:
        w, nr = g["w%d" % i], int(g["n%d" % i])
        ranks, eff = capi.knapsack(w, nr)
        assert ranks.min() >= 0 and ranks.max() < nr
        if len(w) > nr:
            assert abs(eff - float(g["e%d" % i])) < 1e-14
        pa, pb = {}, {}
        for b, (x, y) in enumerate(zip(ranks, g["b%d" % i])):
            pa.setdefault(int(x), set()).add(b)
            pb.setdefault(int(y), set()).add(b)
        assert sorted(map(sorted, pa.values())) == sorted(map(sorted, pb.values())), i
        i += 1


def test_knapsack_vs_port(plm_lib, port_lib):
    from pelelm_b200 import capi
    P = port_lib.Port("LiDryer")
    rng = np.random.default_rng(11)
    for nb, nr in [(1, 1), (3, 2), (64, 2), (200, 8), (513, 4), (9, 8)]:
        w = rng.integers(1, 10 ** 6, nb)
        a, ea = capi.knapsack(w, nr)
        b, eb = P.knapsack(w, nr)
        assert np.array_equal(a, b) and ea == eb, (nb, nr)
    # heaviest bucket goes to the least-loaded rank; empty weights are legal
    a, _ = capi.knapsack(np.zeros(16, dtype=np.int64), 4)
    assert a.min() >= 0 and a.max() < 4


def test_reference_factory_knows_b200_reactors():
    """The adaptor (pelelm_b200/adaptor) registers with the reference's own factory; without a GPU creation must abort
    with the library's no-device message (no silent CPU path)."""
    import subprocess
    import sys
    import torch
    from oracle import pyoracle
    if not pyoracle.ref_available("LiDryer", b200=True):
        pytest.skip("oracle/_ref/libref_LiDryer_b200.so not built")
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by tests/test_gpu_dropin.py")
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from oracle.pyoracle import Ref\n"
            "Ref('LiDryer', b200=True).init(integrator='ReactorRK64B200')\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "Creating ReactorBase instance: ReactorRK64B200" in r.stdout
    assert "no CUDA device" in r.stderr


def test_knapsack_ordered_is_the_sort_true_mapping(plm_lib):
    """plm_knapsack_ordered / plm_least_used_ranks: KnapSackDoIt's bucket -> rank step with the reference's default
    sort = true (AMReX_DistributionMapping.cpp:760-815): the i-th heaviest bucket goes to the i-th least-used rank
    (LeastUsedCPUs :191-235, stable on ties); with at most one box per rank, box i goes to ord[i % nranks]
    (RoundRobinDoIt :365-447).  The partition itself is the one already pinned against AMReX's knapsack()."""
    from pelelm_b200 import capi
    rng = np.random.default_rng(5)
    kb = [700, 100, 100, 900, 50, 100, 800, 0]
    ord8 = capi.least_used_ranks(kb)
    assert ord8.tolist() == [7, 4, 1, 2, 5, 0, 6, 3]          # ascending load, rank order kept among equals
    for nb, nr, order in [(64, 2, np.array([1, 0])), (200, 8, ord8), (513, 4, np.array([2, 0, 3, 1])), (8, 8, ord8), (5, 8, ord8)]:
        w = rng.integers(1, 10 ** 6, nb)
        a, ea = capi.knapsack(w, nr)                           # identity order: bucket i -> rank i (heaviest first)
        b, eb = capi.knapsack(w, nr, rank_order=order)
        assert ea == eb
        assert np.array_equal(b, np.asarray(order)[a])         # same buckets, renamed through the order
        if nb > nr:
            load = np.bincount(b, weights=w, minlength=nr)
            assert np.all(np.diff(load[np.asarray(order)]) <= 0)   # heaviest bucket sits on the least-used rank
    # identity order reproduces plm_knapsack exactly; a non-permutation is rejected
    w = rng.integers(1, 1000, 40)
    assert np.array_equal(capi.knapsack(w, 4)[0], capi.knapsack(w, 4, rank_order=np.arange(4))[0])
    with pytest.raises(capi.PlmError):
        capi.knapsack(w, 4, rank_order=np.array([0, 1, 1, 3]))
    with pytest.raises(capi.PlmError):
        capi.knapsack(w, 4, rank_order=np.array([0, 1, 2]))
