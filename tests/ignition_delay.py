"""The reference's one known-answer test for this path: PelePhysics Testing/Exec/IgnitionDelay (dodecane_lu, 0-D h-reactor,
T0 = 600 K, p = 50 atm, phi = 7.93, ReactorCvode with the difference-quotient Jacobian `dense_direct`, max_order 4,
rtol = atol = 1e-10): ignition delay = first time T exceeds T0 + 400 K, found in two passes (exec_ignDelay.sh,
computeIgnitionDelay.py), accepted by check_ignDelay.py when |tIgn - 0.0763 s| <= 10 * (dt / ndt).

`ignition_delay(react)` restates that driver (utils/reactFunctions.H:20-57: ndt consecutive react() calls of dt/ndt on
one cell; inputs/inputs.0d_firstpass; GPU_misc.H:13-90 for the initial state) around any react(rY, T, rE, dt) -> rY, T, rE, FC.

REFERENCE below = the UNMODIFIED reference (oracle/_ref/libref_dodecane_lu.so) run through this same driver in the
build container (`python tests/ignition_delay.py ref`): the numbers the port and the CUDA path are pinned to.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PUBLISHED = 0.0763          # check_ignDelay.py:23
P0, T0, PHI = 5066250.0, 600.0, 7.928571428571429   # inputs/inputs.0d_firstpass: hr.press, hr.t0, hr.equiv_ratio
INIT = dict(solve_type="dense_direct", rtol=1e-10, atol=1e-10, max_order=4)
REFERENCE = {"estimate": 0.076, "estimate_calls": 76, "t_ign": 0.07634770000000002, "uncertainty": 2.0900000000000003e-05,
             "calls": 3653, "fc_pass1": 8863, "fc_pass2": 234420}


def initial_state():
    from pelelm_b200 import synth
    m = synth.mech("dodecane_lu")
    X = np.zeros(m.ns)
    X[m.idx("O2")] = 1.0 / (PHI / 18.5 + 1 + 0.79 / 0.21)          # GPU_misc.H:56-58
    X[m.idx("N2")] = 0.79 * X[m.idx("O2")] / 0.21
    X[m.idx("NC12H26")] = 1.0 - X[m.idx("O2")] - X[m.idx("N2")]
    Y, rho, h = synth._mixture_state(m, X[None, :].copy(), np.array([T0]), P0)   # X2Y, PYT2R, TY2H
    return (Y[0] * rho[0])[:, None].copy(), np.array([T0]), np.array([rho[0] * h[0]])


def ignition_delay(react):
    """react(rY [ns][1], T [1], rE [1], dt) -> (rY, T, rE, FC).  Returns a dict shaped like REFERENCE."""
    def run(dt, ndt):
        rY, T, rE = initial_state()
        dti, fc = dt / ndt, 0
        for i in range(1, ndt + 1):
            rY, T, rE, FC = react(rY, T, rE, dti)
            fc += int(FC[0])
            if T[0] > T0 + 400.0:      # computeIgnitionDelay.py:98: first row of PPreaction.txt above T0 + 400
                return dti * i, dti, i, fc
        raise RuntimeError("no ignition within dt")
    est, _, i1, fc1 = run(2.0, 2000)               # inputs.0d_firstpass: ode.dt = 2, ode.ndt = 2000
    t, unc, i2, fc2 = run(est * 1.1, 4000)         # computeIgnitionDelay.py:107: dt = 1.1 * estimate, ndt doubled
    return {"estimate": est, "estimate_calls": i1, "t_ign": t, "uncertainty": unc, "calls": i2, "fc_pass1": fc1, "fc_pass2": fc2}


def checker_react(R):
    ns = R.n - 1
    src, esrc, mask = np.zeros((ns, 1)), np.zeros(1), np.ones(1, dtype=np.int32)
    return lambda rY, T, rE, dt: R.react_soa(rY, src, T, rE, esrc, mask, dt)


if __name__ == "__main__":
    from oracle import pyoracle
    if sys.argv[1:] == ["ref"]:
        R = pyoracle.Ref("dodecane_lu").init(integrator="ReactorCvode", **INIT)
    else:
        pyoracle.build_port()
        R = pyoracle.Port("dodecane_lu").init(integrator="ReactorCvode", **INIT)
    R.set_typ_vals(np.full(R.n, -1.0))
    print(ignition_delay(checker_react(R)))
