"""Deterministic synthetic thermochemical states for the tests, smoke() and bench.py.

Host-side INPUT generation only (numpy); the chemistry advance itself never runs here.  Every cell is a
function of (seed, global cell index) alone -- a counter-based splitmix64 stream -- so any rank can
generate any shard of a box and all ranks agree on the global batch.

State recipe per cell, p = 1 atm (1 013 250 dyn/cm^2):
  T      ~ U[900, 2600] K
  phi    ~ logU[0.5, 2]  fuel/air equivalence ratio (fuel H2 | CH4 | NC12H26, air O2:N2 = 0.21:0.79)
  c      ~ U[0, 1]       burnt fraction: the fresh mixture is blended (by mass) with its complete-combustion
                         products (CO2/CO/H2O/H2 split by the oxygen available) plus a radical pool
                         (H, O, OH, HO2 at <= 3e-3 mole fraction) so reaction zones are present and stiff
  rho    = p W / (R T),  rho*h from the NASA-7 tables of pelelm_b200/mechgen/data/<mech>.json
Units are CGS as handed to ReactorBase::react by PeleLM (PeleLM.cpp:5737-5746): rhoY g/cm^3, rho*h erg/cm^3.

`reacteval_state` reproduces the ReactEval initial condition (PelePhysics Testing/Exec/ReactEval/GPU_misc.H:13-82:
X_fuel 0.1, X_O2 0.2, X_N2 0.7, T(y) = 2000 + 500 y/L + 100 sin(2 pi y/(L/4))) used as cfg0.
"""
from __future__ import annotations

import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
RU = 8.31446261815324e7  # erg/mol/K (mechanism.H:145)
P_ATM = 1013250.0
FUEL = {"LiDryer": "H2", "drm19": "CH4", "dodecane_lu": "NC12H26"}
FUEL_CH = {"H2": (0, 2), "CH4": (1, 4), "NC12H26": (12, 26)}
DEFAULT_SEED = 20261017

_cache: dict = {}


class Mech:
    """Species names, molecular weights and NASA-7 h/RT tables of one mechanism."""

    def __init__(self, name: str):
        with open(os.path.join(HERE, "mechgen", "data", name + ".json")) as f:
            d = json.load(f)
        self.name = name
        self.species = d["species"]
        self.ns = len(self.species)
        self.n = self.ns + 1
        self.mw = np.array(d["mw"], dtype=np.float64)
        self.tmid = np.array(d["thermo"]["tmid"], dtype=np.float64)
        self.h = np.array(d["thermo"]["h"], dtype=np.float64)  # [ns][2][6]: a0, a1/2, a2/3, a3/4, a4/5, a5
        self.nreac = d["n_reactions"]

    def idx(self, sp: str) -> int:
        return self.species.index(sp)

    def h_mass(self, T: np.ndarray) -> np.ndarray:
        """Species enthalpies [ncells][ns] in erg/g (CKHMS, mechanism.H:2770-2786)."""
        T = np.asarray(T, dtype=np.float64)[:, None]
        rg = (T >= self.tmid[None, :]).astype(np.int64)
        c = self.h[np.arange(self.ns)[None, :], rg]  # [nc][ns][6]
        hrt = c[..., 0] + c[..., 1] * T + c[..., 2] * T**2 + c[..., 3] * T**3 + c[..., 4] * T**4 + c[..., 5] / T
        return hrt * RU * T / self.mw[None, :]


def mech(name: str) -> Mech:
    if name not in _cache:
        _cache[name] = Mech(name)
    return _cache[name]


def splitmix64(x: np.ndarray) -> np.ndarray:
    """One splitmix64 output per 64-bit input counter (vectorised, wraps modulo 2^64)."""
    with np.errstate(over="ignore"):
        z = x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniforms(seed: int, cell0: int, ncells: int, nstreams: int) -> np.ndarray:
    """[nstreams][ncells] doubles in [0,1): stream s of cell i = splitmix64(splitmix64(seed ^ i) + s)."""
    idx = np.arange(cell0, cell0 + ncells, dtype=np.uint64)
    base = splitmix64(np.uint64(seed) ^ idx)
    out = np.empty((nstreams, ncells))
    with np.errstate(over="ignore"):
        for s in range(nstreams):
            out[s] = (splitmix64(base + np.uint64(s + 1)) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
    return out


def _mixture_state(m: Mech, X: np.ndarray, T: np.ndarray, p: float = P_ATM):
    X = X / X.sum(1, keepdims=True)
    Y = X * m.mw[None, :]
    Y /= Y.sum(1, keepdims=True)
    wbar_inv = (Y / m.mw[None, :]).sum(1)
    rho = p / (RU * T * wbar_inv)
    h = (Y * m.h_mass(T)).sum(1)
    return Y, rho, h


def flame_state(m: Mech, mech_name: str, T: np.ndarray, phi: np.ndarray, c: np.ndarray, urad: np.ndarray,
                p: float = P_ATM):
    """Mass fractions, density and enthalpy (CGS) of fuel/air at equivalence ratio phi, blended with its complete-combustion
    products (+ OH, H, O, HO2 at levels scaled by urad[0..3] in [0,1]) by the burnt fraction c, at temperature T, pressure p."""
    ncells = T.shape[0]
    ns = m.ns
    fuel = FUEL[mech_name]
    x, y = FUEL_CH[fuel]
    a = x + y / 4.0  # O2 per fuel at stoichiometry
    Xf = np.zeros((ncells, ns))
    Xf[:, m.idx(fuel)] = 1.0
    Xf[:, m.idx("O2")] = a / phi
    Xf[:, m.idx("N2")] = a / phi * (0.79 / 0.21)
    # complete-combustion products from the oxygen atoms available
    b = 2.0 * a / phi
    Xb = np.zeros((ncells, ns))
    Xb[:, m.idx("N2")] = a / phi * (0.79 / 0.21)
    o_after_c = b - x  # every carbon becomes at least CO
    h2o = np.clip(o_after_c, 0.0, y / 2.0)
    Xb[:, m.idx("H2O")] = h2o
    Xb[:, m.idx("H2")] += y / 2.0 - h2o
    left = o_after_c - h2o
    if x > 0:
        co2 = np.clip(left, 0.0, x)
        Xb[:, m.idx("CO2")] = co2
        Xb[:, m.idx("CO")] = x - co2
        left = left - co2
    Xb[:, m.idx("O2")] += np.maximum(left, 0.0) / 2.0
    Xb /= Xb.sum(1, keepdims=True)
    for sp, lvl, s in (("OH", 3e-3, 3), ("H", 1e-3, 4), ("O", 1e-3, 5), ("HO2", 1e-5, 6)):
        Xb[:, m.idx(sp)] += lvl * (0.1 + 0.9 * urad[s - 3])
    Yf, _, _ = _mixture_state(m, Xf, T)
    Yb, _, _ = _mixture_state(m, Xb, T)
    Y = (1.0 - c)[:, None] * Yf + c[:, None] * Yb
    wbar_inv = (Y / m.mw[None, :]).sum(1)
    rho = p / (RU * T * wbar_inv)
    h = (Y * m.h_mass(T)).sum(1)
    return Y, rho, h


def synthetic_batch(mech_name: str, ncells: int, seed: int = DEFAULT_SEED, cell0: int = 0, forcing: bool = False,
                    dt: float = 1e-5, mask_every: int = 0, t_range=(900.0, 2600.0)):
    """Component-major (FAB layout) arrays of one batch: dict with rY [ns][nc], rYsrc [ns][nc], T, rE, rEsrc, mask.

    T holds the initial GUESS handed to react (true T + up to +-30 K), as PeleLM hands over a lagged temperature.
    forcing=True adds src_k ~ U[-1,1] 1e-2 rhoY_k/dt with zero net mass and rhoh_src ~ U[-1,1] 1e-2 |rho h|/dt.
    mask_every=k marks every k-th global cell -1 (ReactorCvode skips it; ReactorRK64 ignores the mask).
    """
    m = mech(mech_name)
    ns = m.ns
    u = uniforms(seed, cell0, ncells, 8 + 2 * ns)
    T = t_range[0] + (t_range[1] - t_range[0]) * u[0]
    phi = 0.5 * np.exp(np.log(4.0) * u[1])
    c = u[2]
    Y, rho, h = flame_state(m, mech_name, T, phi, c, u[3:7])
    rY = np.ascontiguousarray((Y * rho[:, None]).T)
    rE = rho * h
    Tguess = T + 60.0 * (u[7] - 0.5)
    rYsrc = np.zeros((ns, ncells))
    rEsrc = np.zeros(ncells)
    if forcing:
        s = (2.0 * u[8:8 + ns] - 1.0) * 1e-2 * rY / dt
        s -= rY / rY.sum(0, keepdims=True) * s.sum(0, keepdims=True)  # zero net mass source
        rYsrc = np.ascontiguousarray(s)
        rEsrc = (2.0 * u[8 + ns] - 1.0) * 1e-2 * np.abs(rE) / dt
    mask = np.ones(ncells, dtype=np.int32)
    if mask_every:
        gidx = np.arange(cell0, cell0 + ncells)
        mask[gidx % mask_every == mask_every - 1] = -1
    return {"rY": rY, "rYsrc": rYsrc, "T": Tguess, "rE": rE, "rEsrc": rEsrc, "mask": mask, "T_true": T}


def reacteval_state(mech_name: str, nx: int, ny: int, nz: int):
    """ReactEval initial condition on an nx*ny*nz box (GPU_misc.H:13-82), component-major, zero forcing."""
    m = mech(mech_name)
    ns = m.ns
    nc = nx * ny * nz
    j = (np.arange(nc) // nx) % ny
    plo, phi_ = -1.0, 1.0
    dx = (phi_ - plo) / ny
    yy = plo + (j + 0.5) * dx
    L = phi_ - plo
    P = L / 4.0
    T = 2000.0 + 500.0 * yy / L + 100.0 * np.sin(2.0 * np.pi * yy / P)
    X = np.zeros((nc, ns))
    X[:, m.idx(FUEL[mech_name])] = 0.1
    X[:, m.idx("O2")] = 0.2
    X[:, m.idx("N2")] = 0.7
    Y, rho, h = _mixture_state(m, X, T)
    return {"rY": np.ascontiguousarray((Y * rho[:, None]).T), "rYsrc": np.zeros((ns, nc)), "T": T.copy(),
            "rE": rho * h, "rEsrc": np.zeros(nc), "mask": np.ones(nc, dtype=np.int32), "T_true": T}
