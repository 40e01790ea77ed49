"""CFL rule of the ns2d.strat time stepper on the GPU path against the oracle (itself pinned to the
reference's TimeSteppingPseudoSpectralStrat, tests/test_oracle.py): advective CFL + internal-wave limits
(solvers/ns2d/strat/time_stepping.py:31-186)."""
import numpy as np
import pytest

from helpers import load_golden, make_gpu_sim, make_oracle, set_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("coef_group", [None, 0.5])
def test_cfl_time_increment_ns2d_strat_matches_oracle(coef_group):
    meta, z = load_golden("ns2d_strat_32x32_rk4")
    meta = dict(meta, params=dict(meta["params"], N=3.0, deltat0=0.2))
    from fluidsim_b200.solvers import SIMUL_CLASSES

    probe = make_gpu_sim(meta, mask=z["mask"])
    p = probe.params
    p.time_stepping.USE_CFL = True
    p.time_stepping.cfl_coef_group = coef_group
    sim = SIMUL_CLASSES["ns2d.strat"](p)
    ts = sim.time_stepping
    o = make_oracle(meta)
    o.cfl_coef_group = coef_group
    lim = o.strat_time_increments()
    assert abs(ts.deltat_dispersion_relation - lim["dispersion_relation"]) < 1e-14 * lim["dispersion_relation"]
    if coef_group:
        assert abs(ts.deltat_group_vel - lim["group_vel"]) < 1e-13 * lim["group_vel"]
        assert abs(ts.deltat_phase_vel - lim["phase_vel"]) < 1e-13 * lim["phase_vel"]
    state = 0.05 * z["state0"]
    for scale in (1.0, 40.0, 41.0, 0.0):
        state = scale * state
        o.set_state_spect(state)
        set_state(sim, state)
        ts.compute_time_increment_CLF()
        want = o.compute_time_increment_CFL(cfl=ts.CFL, deltat_max=ts.deltat_max)
        assert abs(ts.deltat - want) <= 1e-12 * want
    assert ts.CFL == 1.0  # RK4
