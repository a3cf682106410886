"""Development check run on the GPU box: CUDA path vs the two checkers on ReactEval-like states."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.pyoracle import Ref, Port, ref_available
from pelelm_b200.capi import Reactor

def states(P, mech, nc, seed=3, rad=True):
    fuel = {"LiDryer": "H2", "drm19": "CH4", "dodecane_lu": "NC12H26"}[mech]
    rng = np.random.default_rng(seed); sp = P.species; ns = P.ns
    X = np.zeros((nc, ns)); X[:, sp.index(fuel)] = 0.1; X[:, sp.index("O2")] = 0.2; X[:, sp.index("N2")] = 0.7
    if rad:
        X += rng.uniform(0, 1, (nc, ns)) ** 6 * 0.02
    X /= X.sum(1, keepdims=True)
    Y = X * P.mw; Y /= Y.sum(1, keepdims=True)
    T = rng.uniform(900, 2600, nc)
    rho, h = P.pty2rh(1013250.0, T, Y)
    return Y, T, rho, h

def main():
    mechs = sys.argv[1:] or ["drm19", "LiDryer", "dodecane_lu"]
    for mech in mechs:
        P = Port(mech); ns = P.ns; n = ns + 1
        nc = 256
        Y, T, rho, h = states(P, mech, nc)
        rng = np.random.default_rng(5)
        y = np.concatenate([rho[:, None] * Y, (T + rng.uniform(-50, 50, nc))[:, None]], 1)
        rhoe = rho * h; src = rng.uniform(-1, 1, nc) * 1e-2 * np.abs(rhoe) / 1e-5
        rYsrc = rng.uniform(-1, 1, (nc, ns)) * 1e-2 * y[:, :ns] / 1e-5
        G = Reactor(mech, "ReactorCvode").init(rtol=1e-10, atol=1e-10)
        fg = G.eval_rhs(y, rhoe, src, rYsrc, 3e-6); fp = P.rhs(y, rhoe, src, rYsrc, 3e-6)
        print(mech, "RHS gpu vs port: err/max", (np.abs(fg - fp) / np.abs(fp).max(1, keepdims=True)).max(), flush=True)
        y2 = y.copy(); y2[:, ns] = T
        Jg = G.eval_jac(y2); Jp = P.jac(y2)
        print(mech, "JAC gpu vs port: err/max", (np.abs(Jg - Jp) / np.abs(Jp).max(1, keepdims=True)).max(), flush=True)
        # integration
        nci = {"drm19": 128, "LiDryer": 128, "dodecane_lu": 16}[mech]
        Y, T, rho, h = states(P, mech, nci, seed=11, rad=False)
        rY = np.ascontiguousarray((Y * rho[:, None]).T); rE = rho * h
        z = np.zeros_like(rY); ze = np.zeros(nci); mask = np.ones(nci, dtype=np.int32); mask[::9] = -1
        for integ in ["ReactorCvode", "ReactorRK64"]:
            if integ == "ReactorRK64" and mech != "LiDryer":
                continue
            P.init(integrator=integ, rtol=1e-10, atol=1e-10)
            t0 = time.time(); pr = P.react_soa(rY, z, T + 30, rE, ze, mask, 1e-5, nthreads=8); t1 = time.time()
            G = Reactor(mech, integ).init(rtol=1e-10, atol=1e-10)
            a = [rY.copy(), z.copy(), (T + 30).copy(), rE.copy(), ze.copy(), np.zeros(nci)]
            st = G.react([0, 0, 0], [nci - 1, 0, 0], a[0], a[1], a[2], a[3], a[4], a[5], mask, 1e-5)
            print(mech, integ, "stats", st, flush=True)
            rel = np.abs(a[0] - pr[0]) / (np.abs(pr[0]) + 1e-9 * pr[0].sum(0, keepdims=True))
            print(mech, integ, "rhoY rel (floor 1e-9 rho):", rel.max(), "T rel:", np.abs(a[2] / pr[1] - 1).max(),
                  "rE equal:", np.array_equal(a[3], pr[2]), "FC equal frac:", (a[5] == pr[3]).mean(),
                  "FC gpu/port mean", a[5].mean(), pr[3].mean(), "port cells/s", nci / (t1 - t0), flush=True)
            print("   status nonzero:", np.count_nonzero(G.cell_status(nci)))

if __name__ == "__main__":
    main()
