"""Algorithmic FP64 work of one chemistry advance, from the integrator counters.

FLOPs = n_rhs*F_rhs + n_jac*F_jac + n_lu*F_lu + n_solve*F_solve + n_step*F_step, 1 flop per + - * / and ONE flop per
exp/log/log10/pow (useful work, not the instruction expansion).  Constants are static counts of the reference's
generated kernels (productionRate, aJacobian, thermo; one temperature branch, T-Newton counted once; +-10 %).
"""
from __future__ import annotations

#            F_rhs  F_jac   F_lu    F_solve F_step  rk64_step
FLOPS = {
    "LiDryer": (906, 2305, 667, 200, 400, 5796),
    "drm19": (2884, 9197, 7099, 968, 880, 18096),
    "dodecane_lu": (8185, 32506, 104976, 5832, 2160, 51054),
}
NSPEC = {"LiDryer": 9, "drm19": 21, "dodecane_lu": 53}


def state_bytes_per_cell(mech: str) -> int:
    """Algorithmic HBM traffic of the load/store phase: read (2N+3) doubles + mask, write N+3 doubles."""
    n = NSPEC[mech]
    return (2 * n + 3) * 8 + 4 + (n + 3) * 8


def algorithmic_flops(mech: str, integrator: str, st: dict) -> float:
    f_rhs, f_jac, f_lu, f_solve, f_step, f_rk = FLOPS[mech]
    if integrator == "ReactorRK64":
        return float(st["nst"]) * f_rk
    return (float(st["nfe"]) * f_rhs + float(st["nje"]) * f_jac + float(st["nsetups"]) * f_lu +
            float(st["nni"]) * f_solve + float(st["nst"]) * f_step)
