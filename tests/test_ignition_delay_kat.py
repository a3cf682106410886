"""Known-answer test of the reference (PelePhysics Testing/Exec/IgnitionDelay, 0.0763 s) through the oracle port (CPU) and
through the CUDA path (GPU); see tests/ignition_delay.py."""
import numpy as np
import pytest

import ignition_delay as ign


def _check(res, fc_rel):
    ref = ign.REFERENCE
    assert abs(res["t_ign"] - ign.PUBLISHED) <= 10 * res["uncertainty"]          # the reference's own acceptance test
    assert res["estimate_calls"] == ref["estimate_calls"] and res["calls"] == ref["calls"]   # same sub-interval as the reference
    assert res["t_ign"] == ref["t_ign"] and res["uncertainty"] == ref["uncertainty"]
    # FC (RHS evaluations incl. the n per difference-quotient Jacobian): thousands of re-initialised integrations whose
    # Jacobians amplify last-bit differences of exp/log, so only the totals are comparable
    assert abs(res["fc_pass2"] - ref["fc_pass2"]) <= fc_rel * ref["fc_pass2"]


def test_port_reproduces_the_published_ignition_delay(port_lib):
    P = port_lib.Port("dodecane_lu").init(integrator="ReactorCvode", **ign.INIT)
    P.set_typ_vals(np.full(P.n, -1.0))
    _check(ign.ignition_delay(ign.checker_react(P)), 0.01)


@pytest.mark.gpu
def test_cuda_path_reproduces_the_published_ignition_delay(plm_lib):
    from pelelm_b200 import capi
    R = capi.Reactor("dodecane_lu", "ReactorCvode").init(**ign.INIT)
    ns = R.ns
    src, esrc, mask = np.zeros((ns, 1)), np.zeros(1), np.ones(1, dtype=np.int32)

    def react(rY, T, rE, dt):
        rY, T, rE, FC = rY.copy(), T.copy(), rE.copy(), np.zeros(1)
        st = R.react([0, 0, 0], [0, 0, 0], rY, src, T, rE, esrc, FC, mask, dt)
        assert st["nfailed"] == 0
        return rY, T, rE, FC

    res = ign.ignition_delay(react)
    R.close()
    _check(res, 0.05)
