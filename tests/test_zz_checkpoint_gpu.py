"""Checkpoint / restart of the GPU Simul in fluidsim's state_phys layout:
``sim.output.phys_fields.save()`` then ``params.init_fields.type = "from_file"``
(base/output/phys_fields.py:130-202, base/init_fields.py:140-298)."""
import numpy as np
import pytest

from helpers import load_golden, make_gpu_sim, rel_err, set_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["ns3d_16x16x16_rk4", "strat_16x8x32_rk2", "ns2d_32x32_rk4", "ns2d_strat_32x32_rk4"])
def test_restart_from_state_phys_file(name, tmp_path):
    meta, z = load_golden(name)
    sim = make_gpu_sim(meta, mask=z["mask"])
    set_state(sim, z["state0"])
    for _ in range(2):
        sim.time_stepping.one_time_step()
    sim.output.path_run = str(tmp_path)
    path = sim.output.phys_fields.save()

    from fluidsim_b200.minihdf5 import read_hdf5

    root = read_hdf5(path)
    assert root["state_phys"]["@attrs"]["it"] == 2
    assert abs(root["state_phys"]["@attrs"]["time"] - sim.time_stepping.t) < 1e-15
    for key in sim.state.keys_state_phys:
        assert np.array_equal(root["state_phys"][key], sim.state.state_phys.get_var(key).cpu().numpy())

    # restart: a second simulation initialised from the file continues like the first one
    from fluidsim_b200.solvers import SIMUL_CLASSES
    import copy

    p2 = copy.deepcopy(sim.params)
    p2.init_fields.type = "from_file"
    p2.init_fields.from_file.path = path
    sim2 = SIMUL_CLASSES[meta["solver"]](p2)
    sim2.oper.where_dealiased = sim.oper.where_dealiased
    sim2.mask_modified()
    assert sim2.time_stepping.it == 2 and sim2.time_stepping.t == sim.time_stepping.t
    # X -> K of the saved physical fields reproduces the spectral state up to transform round-off
    assert rel_err(sim2.state.state_spect.numpy(), sim.state.state_spect.numpy()) < 1e-13
    for s in (sim, sim2):
        for _ in range(2):
            s.time_stepping.one_time_step()
    assert rel_err(sim2.state.state_spect.numpy(), sim.state.state_spect.numpy()) < 1e-12
