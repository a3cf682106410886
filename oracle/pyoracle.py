"""TEST INFRASTRUCTURE -- ctypes bindings for the two checkers.

* ``Ref(mech)``  : the UNMODIFIED reference compiled into oracle/_ref/libref_<mech>.so (oracle/Makefile, `make ref`).
* ``Port(mech)`` : the plain-C restatement compiled into oracle/_build/liboracle.so (`make port`).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module; nothing under pelelm_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
MAXN = 65


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_dp)


def _pi(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_ip)


def build_port():
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])


def build_ref():
    if os.path.isdir(os.environ.get("PLM_REFERENCE", "/root/reference")):
        subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref"])
        if os.path.exists(os.path.join(os.path.dirname(HERE), "pelelm_b200", "libplmchem.so")):
            subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref_b200"])


def ref_available(mech: str, b200: bool = False) -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref_%s%s.so" % (mech, "_b200" if b200 else "")))


class Opts(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("typ_vals", C.c_double * MAXN),
                ("max_order", C.c_int), ("max_nls_iters", C.c_int), ("clean_init_massfrac", C.c_int),
                ("rk64_guess", C.c_int), ("rk64_min", C.c_int), ("rk64_max", C.c_int),
                ("analytical_jacobian", C.c_int), ("reactor_type", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("nst", C.c_long), ("nfe", C.c_long), ("nje", C.c_long), ("nsetups", C.c_long),
                ("nni", C.c_long), ("ncfn", C.c_long), ("netf", C.c_long), ("flag", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Ref:
    """The reference itself. One mechanism per process-wide library; the reactor is (re)created by ``init``."""

    def __init__(self, mech: str, b200: bool = False):
        # b200=True: the same reference objects + the drop-in adaptor (pelelm_b200/adaptor), so that the reference's
        # own factory can create "ReactorCvodeB200"/"ReactorRK64B200" (oracle/Makefile target ref_b200)
        path = os.path.join(HERE, "_ref", "libref_%s%s.so" % (mech, "_b200" if b200 else ""))
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        self.L = L = C.CDLL(path)
        self.mech = mech
        self.ns = L.ref_num_species()
        self.n = self.ns + 1
        L.ref_reactor_init.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_double, C.c_double] + [C.c_int] * 6
        L.ref_react_box.argtypes = [C.c_int] * 3 + [c_dp] * 6 + [c_ip, C.c_double, C.c_double]
        L.ref_react_flat.argtypes = [C.c_int] + [c_dp] * 4 + [C.c_double, C.c_double]
        L.ref_rhs.argtypes = [C.c_int, C.c_int, C.c_double] + [c_dp] * 5
        L.ref_jac.argtypes = [C.c_int, C.c_int, c_dp, c_dp]
        L.ref_ckwc.argtypes = [C.c_int, c_dp, c_dp, c_dp]
        L.ref_thermo.argtypes = [C.c_int] + [c_dp] * 5
        L.ref_pty2rh.argtypes = [C.c_int, C.c_double] + [c_dp] * 4
        L.ref_set_typ_vals.argtypes = [c_dp]
        L.ref_reaction_rate_rhoY.argtypes = [C.c_int] * 3 + [c_dp] * 3 + [c_ip, c_dp]
        buf = C.create_string_buffer(8192)
        L.ref_species_names(buf, 8192)
        self.species = buf.value.decode().split()

    def init(self, integrator="ReactorCvode", solve_type="denseAJ_direct", rtol=1e-10, atol=1e-10, max_order=2,
             max_nls_iters=3, clean_init_massfrac=0, rk64=(10, 5, 500), reactor_type=2):
        r = self.L.ref_reactor_init(integrator.encode(), solve_type.encode(), reactor_type, rtol, atol, max_order,
                                    max_nls_iters, clean_init_massfrac, *rk64)
        assert r == 0
        return self

    def set_typ_vals(self, typ):
        typ = np.ascontiguousarray(typ, dtype=np.float64)
        assert typ.size == self.n
        self.L.ref_set_typ_vals(_p(typ))

    def react_soa(self, rY, rYsrc, T, rE, rEsrc, mask, dt, time=0.0):
        """In-place on copies; arrays are component-major [ns][ncells]. Returns (rY, T, rE, FC)."""
        nc = T.size
        rY, T, rE = rY.copy(), T.copy(), rE.copy()
        FC = np.zeros(nc)
        mask = np.ascontiguousarray(mask, dtype=np.int32)
        self.L.ref_react_box(nc, 1, 1, _p(rY), _p(np.ascontiguousarray(rYsrc)), _p(T), _p(rE),
                             _p(np.ascontiguousarray(rEsrc)), _p(FC), _pi(mask), dt, time)
        return rY, T, rE, FC

    def reaction_rate_rhoY(self, rhoY, rhoH, T, mask):
        """The reference's own reactionRateRhoY (Source/PeleLM_K.H:57-101) over an nc x 1 x 1 box; SoA [ns][nc], MKS."""
        nc = T.size
        out = np.zeros_like(rhoY)
        self.L.ref_reaction_rate_rhoY(nc, 1, 1, _p(rhoY), _p(rhoH), _p(T), _pi(mask), _p(out))
        return out

    def rhs(self, y, rhoe, rhoesrc, rYsrc, t=0.0, reactor_type=2):
        nc = y.shape[0]
        out = np.zeros_like(y)
        self.L.ref_rhs(nc, reactor_type, t, _p(y), _p(rhoe), _p(rhoesrc), _p(rYsrc), _p(out))
        return out

    def jac(self, y, reactor_type=2):
        nc = y.shape[0]
        J = np.zeros((nc, self.n * self.n))
        self.L.ref_setup()
        self.L.ref_jac(nc, reactor_type, _p(y), _p(J))
        return J

    def ckwc(self, T, Cc):
        w = np.zeros_like(Cc)
        self.L.ref_ckwc(T.size, _p(T), _p(Cc), _p(w))
        return w

    def thermo(self, h, Y, Tguess):
        nc = h.size
        T = Tguess.copy()
        hi = np.zeros((nc, self.ns))
        cp = np.zeros(nc)
        self.L.ref_thermo(nc, _p(h), _p(Y), _p(T), _p(hi), _p(cp))
        return T, hi, cp

    def pty2rh(self, p, T, Y):
        nc = T.size
        rho, h = np.zeros(nc), np.zeros(nc)
        self.L.ref_pty2rh(nc, p, _p(T), _p(Y), _p(rho), _p(h))
        return rho, h


def advance_chemistry_soa(R, rhoY, rhoH, T, frc_rhoY, frc_rhoH, mask, dt, **kw):
    """The body of PeleLM::advance_chemistry around a checker's react (R = Ref or Port), all arrays MKS, SoA [ns][nc]:
    MKS -> CGS (PeleLM.cpp:5737-5746), react (:5752), CGS -> MKS (:5765-5772), I_R = -(old - new)/dt - F (:5792-5799).
    Returns rhoY_new, rhoH_new, T_new, FC, I_R."""
    rY = rhoY * 1.0e-3
    fY = frc_rhoY * 1.0e-3
    rH = rhoH * 10.0
    fH = frc_rhoH * 10.0
    rY2, T2, rE2, FC = R.react_soa(rY, fY, T, rH, fH, mask, dt, **kw)
    rY_new = rY2 * 1.0e3
    rH_new = rE2 * 0.1
    dt_inv = 1.0 / dt
    I_R = -(rhoY - rY_new) * dt_inv - frc_rhoY
    return rY_new, rH_new, T2, FC, I_R


class Port:
    """The plain-C restatement (all three mechanisms in one library)."""
    _lib = None

    def __init__(self, mech: str, build: bool = True):
        if Port._lib is None:
            path = os.path.join(HERE, "_build", "liboracle.so")
            if build:
                build_port()
            Port._lib = C.CDLL(path)
            L = Port._lib
            L.orc_mech.restype = C.c_void_p
            L.orc_mech.argtypes = [C.c_char_p]
            L.orc_react_soa.argtypes = [C.c_char_p, C.POINTER(Opts), C.c_int, C.c_int] + [c_dp] * 6 + \
                [c_ip, C.c_double, C.c_double, C.c_int, C.POINTER(Stats)]
            L.orc_rhs.argtypes = [C.c_void_p, C.c_double, c_dp, c_dp, C.c_double, C.c_double, c_dp]
            L.orc_cjac.argtypes = [C.c_void_p, c_dp, c_dp]
            L.orc_ckwc.argtypes = [C.c_void_p, C.c_double, c_dp, c_dp]
            L.orc_get_T_given_hY.argtypes = [C.c_void_p, C.c_double, c_dp, c_dp]
            L.orc_hms.argtypes = [C.c_void_p, C.c_double, c_dp]
            L.orc_cpbs.argtypes = [C.c_void_p, C.c_double, c_dp]
            L.orc_cpbs.restype = C.c_double
            L.orc_hbms.argtypes = [C.c_void_p, C.c_double, c_dp]
            L.orc_hbms.restype = C.c_double
            L.orc_default_opts.argtypes = [C.POINTER(Opts)]
            L.orc_knapsack.argtypes = [C.POINTER(C.c_long), C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, c_ip, c_dp]
        self.L = Port._lib
        self.mech = mech
        self.m = self.L.orc_mech(mech.encode())
        assert self.m, mech
        import json
        with open(os.path.join(os.path.dirname(HERE), "pelelm_b200", "mechgen", "data", mech + ".json")) as f:
            d = json.load(f)
        self.species = d["species"]
        self.mw = np.array(d["mw"])
        self.ns = len(self.species)
        self.n = self.ns + 1
        self.opts = Opts()
        self.L.orc_default_opts(C.byref(self.opts))

    def init(self, integrator="ReactorCvode", solve_type="denseAJ_direct", rtol=1e-10, atol=1e-10, max_order=2,
             max_nls_iters=3, clean_init_massfrac=0, rk64=(10, 5, 500), reactor_type=2):
        assert reactor_type in (1, 2)
        o = self.opts
        o.reactor_type = reactor_type
        o.rtol, o.atol, o.max_order, o.max_nls_iters = rtol, atol, max_order, max_nls_iters
        o.clean_init_massfrac = clean_init_massfrac
        o.rk64_guess, o.rk64_min, o.rk64_max = rk64
        o.analytical_jacobian = {"denseAJ_direct": 1, "dense_direct": 0}[solve_type]
        self.use_rk64 = {"ReactorCvode": 0, "ReactorRK64": 1}[integrator]
        return self

    def set_typ_vals(self, typ):
        for i, v in enumerate(typ):
            self.opts.typ_vals[i] = float(v)

    def react_soa(self, rY, rYsrc, T, rE, rEsrc, mask, dt, time=0.0, nthreads=1, return_stats=False):
        nc = T.size
        rY, T, rE = rY.copy(), T.copy(), rE.copy()
        FC = np.zeros(nc)
        mask = np.ascontiguousarray(mask, dtype=np.int32)
        st = Stats()
        self.L.orc_react_soa(self.mech.encode(), C.byref(self.opts), self.use_rk64, nc, _p(rY),
                             _p(np.ascontiguousarray(rYsrc)), _p(T), _p(rE), _p(np.ascontiguousarray(rEsrc)),
                             _p(FC), _pi(mask), dt, time, nthreads, C.byref(st))
        if return_stats:
            return rY, T, rE, FC, st.as_dict()
        return rY, T, rE, FC

    def reaction_rate_rhoY(self, rhoY, rhoH, T, mask):
        nc = T.size
        out = np.zeros_like(rhoY)
        col = np.zeros(self.ns)
        self.L.orc_reaction_rate_rhoY.argtypes = [C.c_void_p, c_dp, C.c_double, C.c_double, C.c_int, c_dp]
        for c in range(nc):
            self.L.orc_reaction_rate_rhoY(self.m, _p(np.ascontiguousarray(rhoY[:, c])), float(rhoH[c]), float(T[c]),
                                          int(mask[c]), _p(col))
            out[:, c] = col
        return out

    def rhs(self, y, rhoe, rhoesrc, rYsrc, t=0.0, reactor_type=2):
        out = np.zeros_like(y)
        self.L.orc_rhs_rt.argtypes = [C.c_void_p, C.c_int, C.c_double, c_dp, c_dp, C.c_double, C.c_double, c_dp]
        for c in range(y.shape[0]):
            self.L.orc_rhs_rt(self.m, reactor_type, t, _p(y[c]), _p(out[c]), rhoe[c], rhoesrc[c], _p(rYsrc[c]))
        return out

    def jac(self, y, reactor_type=2):
        J = np.zeros((y.shape[0], self.n * self.n))
        self.L.orc_cjac_rt.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp]
        for c in range(y.shape[0]):
            self.L.orc_cjac_rt(self.m, reactor_type, _p(y[c]), _p(J[c]))
        return J

    def ckwc(self, T, Cc):
        w = np.zeros_like(Cc)
        for c in range(T.size):
            self.L.orc_ckwc(self.m, T[c], _p(Cc[c]), _p(w[c]))
        return w

    def thermo(self, h, Y, Tguess):
        nc = h.size
        T = Tguess.copy()
        hi = np.zeros((nc, self.ns))
        cp = np.zeros(nc)
        for c in range(nc):
            t = C.c_double(T[c])
            self.L.orc_get_T_given_hY(self.m, h[c], _p(Y[c]), C.cast(C.byref(t), c_dp))
            T[c] = t.value
            self.L.orc_hms(self.m, T[c], _p(hi[c]))
            cp[c] = self.L.orc_cpbs(self.m, T[c], _p(Y[c]))
        return T, hi, cp

    def pty2rh(self, p, T, Y):
        """rho from the ideal-gas law and h = sum Y_k h_k (Fuego.H PYT2R / TY2H)."""
        RU = 8.31446261815324e7
        nc = T.size
        rho = np.zeros(nc)
        h = np.zeros(nc)
        for c in range(nc):
            wbar_inv = float(np.sum(Y[c] / self.mw))
            rho[c] = p / (RU * T[c] * wbar_inv)
            h[c] = self.L.orc_hbms(self.m, T[c], _p(Y[c]))
        return rho, h

    def knapsack(self, wgts, nprocs, max_eff=0.9, do_full=True, nmax=2 ** 31 - 1):
        w = (C.c_long * len(wgts))(*[int(x) for x in wgts])
        out = np.zeros(len(wgts), dtype=np.int32)
        eff = C.c_double(0)
        self.L.orc_knapsack(w, len(wgts), nprocs, max_eff, int(do_full), nmax, _pi(out), C.cast(C.byref(eff), c_dp))
        return out, eff.value
