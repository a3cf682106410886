"""ctypes binding of libplmchem.so (include/plm_chem.h) -- the same C ABI a C++/Fortran caller links.

`Reactor` mirrors the reference operator (PelePhysics/Reactions/ReactorBase.H:15-98): create by name, `init`,
`set_typ_vals_ode`, `react(box, rY, rYsrc, T, rEner, rEner_src, FC, mask, dt_react, time)`, `close`.
Arrays may be numpy arrays (host: staged by the library) or torch CUDA tensors (device pointers are passed
through unchanged). There is no CPU fallback: loading fails loudly if the library or a B200-class GPU is missing.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get("PLM_LIBPATH") or os.path.join(HERE, "libplmchem.so")  # PLM_LIBPATH: development variants

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class A4D(C.Structure):
    _fields_ = [("p", c_dp), ("jstride", C.c_int64), ("kstride", C.c_int64), ("nstride", C.c_int64),
                ("lo", C.c_int * 3), ("hi", C.c_int * 3), ("ncomp", C.c_int)]


class A4I(C.Structure):
    _fields_ = [("p", c_ip), ("jstride", C.c_int64), ("kstride", C.c_int64), ("nstride", C.c_int64),
                ("lo", C.c_int * 3), ("hi", C.c_int * 3), ("ncomp", C.c_int)]


class Box(C.Structure):
    _fields_ = [("lo", C.c_int * 3), ("hi", C.c_int * 3), ("rY", A4D), ("rYsrc", A4D), ("T", A4D), ("rEner", A4D),
                ("rEner_src", A4D), ("FC", A4D), ("I_R", A4D), ("mask", A4I)]


class Stats(C.Structure):
    _fields_ = [("ncells", C.c_int64), ("nskipped", C.c_int64), ("nst", C.c_int64), ("nfe", C.c_int64),
                ("nje", C.c_int64), ("nsetups", C.c_int64), ("nni", C.c_int64), ("ncfn", C.c_int64),
                ("netf", C.c_int64), ("nfailed", C.c_int64), ("max_fc", C.c_int64), ("min_fc", C.c_int64),
                ("kernel_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = ["plm_reactor_create", "plm_reactor_init", "plm_reactor_set_typ_vals", "plm_react_box", "plm_react_flat",
           "plm_advance_chemistry_box", "plm_reaction_rate_rhoY",
           "plm_reactor_stats", "plm_reactor_cell_status", "plm_reactor_close", "plm_num_species", "plm_last_error",
           "plm_eval_rhs", "plm_eval_jac", "plm_box_costs", "plm_knapsack", "plm_knapsack_ordered", "plm_least_used_ranks",
           "plm_measure_fp64_peak", "plm_species_name", "plm_reactor_set_path", "plm_reactor_failed_cells", "plm_react_boxes", "plm_reactor_uses_lane"]

_lib = None


def load():
    """Load libplmchem.so and declare the prototypes. Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise RuntimeError("libplmchem.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    L = C.CDLL(LIBPATH)
    vp = C.c_void_p
    L.plm_reactor_create.argtypes = [C.POINTER(vp), C.c_char_p, C.c_char_p]
    L.plm_reactor_init.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_char_p,
                                   C.c_int, C.c_int, C.c_int]
    L.plm_reactor_set_typ_vals.argtypes = [vp, c_dp, C.c_int]
    L.plm_react_box.argtypes = [vp, c_ip, c_ip, A4D, A4D, A4D, A4D, A4D, A4D, A4I, c_dp, c_dp, vp]
    L.plm_advance_chemistry_box.argtypes = [vp, c_ip, c_ip, A4D, A4D, A4D, A4D, A4D, A4D, A4D, A4I, C.c_double, vp]
    L.plm_reaction_rate_rhoY.argtypes = [vp, c_ip, c_ip, A4D, A4D, A4D, A4I, A4D, vp]
    L.plm_react_flat.argtypes = [vp, vp, vp, vp, vp, c_dp, c_dp, C.c_int, vp]
    L.plm_reactor_stats.argtypes = [vp, C.POINTER(Stats)]
    L.plm_reactor_cell_status.argtypes = [vp, vp, C.c_int64]
    L.plm_reactor_close.argtypes = [vp]
    L.plm_reactor_close.restype = None
    L.plm_num_species.argtypes = [vp]
    L.plm_last_error.restype = C.c_char_p
    L.plm_species_name.argtypes = [vp, C.c_int]
    L.plm_species_name.restype = C.c_char_p
    L.plm_eval_rhs.argtypes = [vp, C.c_int, C.c_double, vp, vp, vp, vp, vp, vp]
    L.plm_eval_jac.argtypes = [vp, C.c_int, vp, vp, vp]
    L.plm_box_costs.argtypes = [vp, vp, C.c_int, vp, vp]
    L.plm_knapsack.argtypes = [C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, c_ip, c_dp]
    L.plm_knapsack_ordered.argtypes = [C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, c_ip, c_ip, c_dp]
    L.plm_least_used_ranks.argtypes = [C.POINTER(C.c_int64), C.c_int, c_ip]
    L.plm_measure_fp64_peak.argtypes = [c_dp, vp]
    L.plm_reactor_set_path.argtypes = [vp, C.c_int]
    L.plm_reactor_uses_lane.argtypes = [vp]
    L.plm_react_boxes.argtypes = [vp, C.c_int, C.POINTER(Box), C.c_double, C.c_double, C.c_int, vp]
    L.plm_reactor_failed_cells.argtypes = [vp, C.POINTER(C.c_longlong)]
    _lib = L
    return L


class PlmError(RuntimeError):
    pass


def _ptr(x):
    """Raw address of a numpy array or torch tensor (float64/int32, contiguous)."""
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    return x.data_ptr()  # torch tensor


def knapsack(wgts, nranks, max_efficiency=0.9, do_full=True, nmax=2 ** 31 - 1, rank_order=None):
    """DistributionMapping::KnapSackProcessorMap on a cost vector (host only; needs no GPU). rank_order (a permutation of
    the ranks, e.g. least_used_ranks(...)) is the reference's sort = true: the i-th heaviest bucket goes to rank_order[i]."""
    L = load()
    w = (C.c_int64 * len(wgts))(*[int(v) for v in wgts])
    out = np.zeros(len(wgts), dtype=np.int32)
    eff = C.c_double(0)
    if rank_order is None:
        rc = L.plm_knapsack(w, len(wgts), nranks, max_efficiency, int(do_full), nmax,
                            out.ctypes.data_as(c_ip), C.cast(C.byref(eff), c_dp))
    else:
        ro = np.ascontiguousarray(rank_order, dtype=np.int32)
        if ro.size != nranks:
            raise PlmError("rank_order must list every rank once")
        rc = L.plm_knapsack_ordered(w, len(wgts), nranks, max_efficiency, int(do_full), nmax, ro.ctypes.data_as(c_ip),
                                    out.ctypes.data_as(c_ip), C.cast(C.byref(eff), c_dp))
    if rc != 0:
        raise PlmError("plm_knapsack failed")
    return out, eff.value


def least_used_ranks(kbytes):
    """DistributionMapping::LeastUsedCPUs: ranks by ascending kilobytes held (stable)."""
    L = load()
    n = len(kbytes)
    kb = (C.c_int64 * n)(*[int(v) for v in kbytes])
    out = np.zeros(n, dtype=np.int32)
    if L.plm_least_used_ranks(kb, n, out.ctypes.data_as(c_ip)) != 0:
        raise PlmError("plm_least_used_ranks failed")
    return out


def measure_fp64_peak() -> float:
    """Register-resident DFMA throughput of the current device, TFLOP/s."""
    L = load()
    v = C.c_double(0)
    rc = L.plm_measure_fp64_peak(C.cast(C.byref(v), c_dp), None)
    if rc != 0:
        raise PlmError(L.plm_last_error().decode())
    return v.value


def box_costs(FC, box_offsets, stream=None):
    """Per-box sum of FC on the device (PeleLM::getFuncCountDM, PeleLM.cpp:5557-5571). torch CUDA tensors."""
    import torch
    L = load()
    nb = box_offsets.numel() - 1
    costs = torch.zeros(nb, dtype=torch.int64, device=FC.device)
    rc = L.plm_box_costs(FC.data_ptr(), box_offsets.data_ptr(), nb, costs.data_ptr(), C.c_void_p(stream or 0))
    if rc != 0:
        raise PlmError(L.plm_last_error().decode())
    return costs


class Reactor:
    """ReactorBase::create(name) + the virtual interface, forwarding to the C ABI."""

    def __init__(self, mechanism: str, chem_integrator: str = "ReactorCvode"):
        self.L = load()
        self.h = C.c_void_p()
        rc = self.L.plm_reactor_create(C.byref(self.h), mechanism.encode(), chem_integrator.encode())
        if rc != 0:
            raise PlmError(self.L.plm_last_error().decode())
        self.mechanism = mechanism
        self.integrator = chem_integrator
        self.ns = self.L.plm_num_species(self.h)
        self.n = self.ns + 1

    def _chk(self, rc):
        if rc != 0:
            raise PlmError("%d: %s" % (rc, self.L.plm_last_error().decode()))

    def init(self, reactor_type=2, rtol=1e-6, atol=1e-10, max_order=2, max_nls_iters=3, clean_init_massfrac=0,
             solve_type="denseAJ_direct", rk64=(10, 5, 500)):
        self._chk(self.L.plm_reactor_init(self.h, reactor_type, rtol, atol, max_order, max_nls_iters,
                                          clean_init_massfrac, solve_type.encode(), *rk64))
        return self

    def path_is_per_box(self):
        """True when a react_boxes call is a loop of per-box launches, False when it is one launch over the cells of every box
        (the default for both kernel families; PLM_BOXES_PER_LAUNCH=1 keeps the warp-per-cell kernels on one launch per box)"""
        if bool(self.L.plm_reactor_uses_lane(self.h)):
            return False
        return os.environ.get("PLM_BOXES_PER_LAUNCH", "") == "1" or (
            os.environ.get("PLM_RK64_DUAL", "") == "1" and self.integrator == "ReactorRK64")

    def set_path(self, path):
        """kernel family: 0 / "auto", 1 / "warp" (one warp per cell), 2 / "lane" (one thread per cell)"""
        path = {"auto": 0, "warp": 1, "lane": 2}.get(path, path)
        self._chk(self.L.plm_reactor_set_path(self.h, int(path)))
        return self

    def failed_cells(self):
        """cells whose integration has failed over the life of the reactor (never blocks)"""
        n = C.c_longlong(0)
        self._chk(self.L.plm_reactor_failed_cells(self.h, C.byref(n)))
        return int(n.value)

    def set_typ_vals_ode(self, typ):
        typ = np.ascontiguousarray(typ, dtype=np.float64)
        self._chk(self.L.plm_reactor_set_typ_vals(self.h, typ.ctypes.data_as(c_dp), typ.size))

    @staticmethod
    def _a4(cls, x, lo, hi, ncomp):
        nx, ny, nz = hi[0] - lo[0] + 1, hi[1] - lo[1] + 1, hi[2] - lo[2] + 1
        a = cls()
        a.p = C.cast(_ptr(x), c_dp if cls is A4D else c_ip)
        a.jstride, a.kstride, a.nstride = nx, nx * ny, nx * ny * nz
        a.lo[:] = lo
        a.hi[:] = hi
        a.ncomp = ncomp
        return a

    def react(self, box_lo, box_hi, rY, rYsrc, T, rEner, rEner_src, FC, mask, dt_react, time=0.0, fab_lo=None,
              fab_hi=None, stream=None, sync=True):
        """In-place on the given arrays (FAB layout: component-major over the FAB box [fab_lo, fab_hi]).

        With device arrays the call only enqueues work on `stream`; sync=True (default) then waits and returns the
        statistics of the call, sync=False returns None -- collect the totals later with stats()."""
        fab_lo = list(box_lo if fab_lo is None else fab_lo)
        fab_hi = list(box_hi if fab_hi is None else fab_hi)
        ns = self.ns
        lo = (C.c_int * 3)(*box_lo)
        hi = (C.c_int * 3)(*box_hi)
        a = [self._a4(A4D, rY, fab_lo, fab_hi, ns), self._a4(A4D, rYsrc, fab_lo, fab_hi, ns),
             self._a4(A4D, T, fab_lo, fab_hi, 1), self._a4(A4D, rEner, fab_lo, fab_hi, 1),
             self._a4(A4D, rEner_src, fab_lo, fab_hi, 1), self._a4(A4D, FC, fab_lo, fab_hi, 1)]
        if mask is None:
            m = A4I()
        else:
            m = self._a4(A4I, mask, fab_lo, fab_hi, 1)
        dt = C.c_double(dt_react)
        tm = C.c_double(time)
        self._chk(self.L.plm_react_box(self.h, lo, hi, *a, m, C.cast(C.byref(dt), c_dp), C.cast(C.byref(tm), c_dp),
                                       C.c_void_p(stream or 0)))
        return self.stats() if sync else None

    def advance_chemistry(self, box_lo, box_hi, rhoY, rhoH, T, frc_rhoY, frc_rhoH, I_R, FC, mask, dt, fab_lo=None,
                          fab_hi=None, stream=None, sync=True):
        """The fused body of PeleLM::advance_chemistry (PeleLM.cpp:5736-5800): MKS arrays; rhoY, rhoH, T, FC updated in
        place, I_R written (pass None to skip), forcing read only."""
        fab_lo = list(box_lo if fab_lo is None else fab_lo)
        fab_hi = list(box_hi if fab_hi is None else fab_hi)
        ns = self.ns
        lo = (C.c_int * 3)(*box_lo)
        hi = (C.c_int * 3)(*box_hi)
        ir = A4D() if I_R is None else self._a4(A4D, I_R, fab_lo, fab_hi, ns)
        m = A4I() if mask is None else self._a4(A4I, mask, fab_lo, fab_hi, 1)
        self._chk(self.L.plm_advance_chemistry_box(
            self.h, lo, hi, self._a4(A4D, rhoY, fab_lo, fab_hi, ns), self._a4(A4D, rhoH, fab_lo, fab_hi, 1),
            self._a4(A4D, T, fab_lo, fab_hi, 1), self._a4(A4D, frc_rhoY, fab_lo, fab_hi, ns),
            self._a4(A4D, frc_rhoH, fab_lo, fab_hi, 1), ir, self._a4(A4D, FC, fab_lo, fab_hi, 1), m, float(dt),
            C.c_void_p(stream or 0)))
        return self.stats() if sync else None

    def react_boxes(self, boxes, dt, time=0.0, mks=False, stream=None, sync=True):
        """One launch over several boxes (device arrays): `boxes` = dicts with lo, hi, rY, rYsrc, T, rEner, rEner_src, FC and
        optionally mask, I_R, fab_lo, fab_hi.  mks=True: the fused advance_chemistry body (MKS arrays, I_R written)."""
        ns = self.ns
        arr = (Box * len(boxes))()
        for e, b in zip(arr, boxes):
            flo, fhi = list(b.get("fab_lo", b["lo"])), list(b.get("fab_hi", b["hi"]))
            e.lo[:] = list(b["lo"])
            e.hi[:] = list(b["hi"])
            e.rY, e.rYsrc = self._a4(A4D, b["rY"], flo, fhi, ns), self._a4(A4D, b["rYsrc"], flo, fhi, ns)
            e.T, e.rEner = self._a4(A4D, b["T"], flo, fhi, 1), self._a4(A4D, b["rEner"], flo, fhi, 1)
            e.rEner_src, e.FC = self._a4(A4D, b["rEner_src"], flo, fhi, 1), self._a4(A4D, b["FC"], flo, fhi, 1)
            e.I_R = self._a4(A4D, b["I_R"], flo, fhi, ns) if b.get("I_R") is not None else A4D()
            e.mask = self._a4(A4I, b["mask"], flo, fhi, 1) if b.get("mask") is not None else A4I()
        self._chk(self.L.plm_react_boxes(self.h, len(boxes), arr, float(dt), float(time), 1 if mks else 0, C.c_void_p(stream or 0)))
        return self.stats() if sync else None

    def reaction_rate_rhoY(self, box_lo, box_hi, rhoY, rhoH, T, mask, rhoYdot, fab_lo=None, fab_hi=None, stream=None):
        """reactionRateRhoY over a box (PeleLM_K.H:57-101): MKS in, rhoYdot (N comps, kg/m^3/s) out."""
        fab_lo = list(box_lo if fab_lo is None else fab_lo)
        fab_hi = list(box_hi if fab_hi is None else fab_hi)
        ns = self.ns
        lo = (C.c_int * 3)(*box_lo)
        hi = (C.c_int * 3)(*box_hi)
        m = A4I() if mask is None else self._a4(A4I, mask, fab_lo, fab_hi, 1)
        self._chk(self.L.plm_reaction_rate_rhoY(
            self.h, lo, hi, self._a4(A4D, rhoY, fab_lo, fab_hi, ns), self._a4(A4D, rhoH, fab_lo, fab_hi, 1),
            self._a4(A4D, T, fab_lo, fab_hi, 1), m, self._a4(A4D, rhoYdot, fab_lo, fab_hi, ns), C.c_void_p(stream or 0)))

    def react_flat(self, rY, rYsrc, rX, rX_src, dt_react, time=0.0, stream=None):
        ncells = int(np.prod(rX.shape))
        dt = C.c_double(dt_react)
        tm = C.c_double(time)
        self._chk(self.L.plm_react_flat(self.h, _ptr(rY), _ptr(rYsrc), _ptr(rX), _ptr(rX_src),
                                        C.cast(C.byref(dt), c_dp), C.cast(C.byref(tm), c_dp), ncells,
                                        C.c_void_p(stream or 0)))
        return self.stats()

    def stats(self):
        """Waits for outstanding calls; totals over the react calls since the previous stats()."""
        s = Stats()
        self._chk(self.L.plm_reactor_stats(self.h, C.byref(s)))
        return s.as_dict()

    def cell_status(self, ncells):
        out = np.zeros(ncells, dtype=np.int32)
        self._chk(self.L.plm_reactor_cell_status(self.h, out.ctypes.data, ncells))
        return out

    def eval_rhs(self, y, rhoe, rhoesrc, rYsrc, t=0.0):
        ydot = np.zeros_like(y) if isinstance(y, np.ndarray) else y.new_zeros(y.shape)
        self._chk(self.L.plm_eval_rhs(self.h, y.shape[0], t, _ptr(y), _ptr(rhoe), _ptr(rhoesrc), _ptr(rYsrc),
                                      _ptr(ydot), None))
        return ydot

    def eval_jac(self, y):
        n = self.n
        J = np.zeros((y.shape[0], n * n)) if isinstance(y, np.ndarray) else y.new_zeros((y.shape[0], n * n))
        self._chk(self.L.plm_eval_jac(self.h, y.shape[0], _ptr(y), _ptr(J), None))
        return J

    def close(self):
        if self.h:
            self.L.plm_reactor_close(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
