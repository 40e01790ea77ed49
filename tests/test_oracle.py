"""CPU tests: the numpy oracle against the committed golden vectors and (in the build container)
against the reference's own modules executed through ``oracle.refshim``."""
import os

import numpy as np
import pytest

from helpers import all_golden_cases, golden_cases, load_golden, make_oracle, rel_err
from oracle import refshim, step_np


@pytest.mark.parametrize("name", all_golden_cases())
def test_oracle_matches_golden(name):
    meta, z = load_golden(name)
    o = make_oracle(meta, stepping=True)  # every scheme of the reference's stepper is restated
    assert np.array_equal(o.oper.where_dealiased, z["mask"])
    o.set_state_spect(z["state0"])
    if "forcing" in z.files:  # forced goldens: tendencies += forcing_fft (solvers/ns3d/solver.py:243-244)
        o.forcing_fft = z["forcing"]
    tend = np.array(o.tendencies_nonlin())
    assert rel_err(tend, z["tend0"]) < 1e-13
    o.one_time_step()
    assert rel_err(o.state_spect, z["state1"]) < 1e-13
    for _ in range(meta["nsteps"] - 1):
        o.one_time_step()
    assert rel_err(o.state_spect, z["stateN"]) < 1e-12
    assert abs(o.compute_energy() - float(z["energyN"])) <= 1e-12 * abs(float(z["energyN"]))


def test_noise_init_is_deterministic_and_solenoidal():
    o = step_np.OracleSim("ns3d", 16, 12, 8, nu_2=1e-2)
    o.init_noise()
    s = np.array(o.state_spect)
    o2 = step_np.OracleSim("ns3d", 16, 12, 8, nu_2=1e-2)
    o2.init_noise()
    assert np.array_equal(s, np.array(o2.state_spect))
    div = o.oper.divfft_from_vecfft(*s)
    assert np.abs(div).max() < 1e-14
    vmax = np.sqrt((np.array(o.state_phys) ** 2).sum(0)).max()
    assert abs(vmax - 1.0) < 1e-12


def test_nonlinear_term_conserves_energy():
    """solvers/ns3d/test_solver.py:73-96 on the oracle."""
    o = step_np.OracleSim("ns3d", 20, 15, 10, nu_2=0.0, coef_dealiasing=2 / 3)
    o.init_noise()
    T = np.array(o.tendencies_nonlin())
    S = np.array(o.state_spect)
    ratio = np.real(T.conj() * S)
    tot = sum(o.oper.sum_wavenumbers(ratio[i]) for i in range(3))
    tot_abs = sum(o.oper.sum_wavenumbers(np.abs(ratio[i])) for i in range(3))
    assert abs(tot) / tot_abs < 1e-13


def test_taylor_green_energy():
    o = step_np.OracleSim("ns3d", 16, 16, 16, nu_2=1 / 1600.0)
    o.init_taylor_green()
    assert abs(o.compute_energy() - 0.125) < 1e-14


def with_buoyancy_2d(o, solver, nx, ny, kw):
    """State of a 2-D buoyancy solver whose b field is a second (dealiased) noise realisation."""
    o2 = step_np.OracleSim(solver, nx, ny, None, **kw)
    o2.init_noise(seed=7)
    s = np.array(o.state_spect)
    s[1] = 0.5 * np.array(o2.state_spect)[0]
    return s


@pytest.mark.skipif(not refshim.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize(
    "solver,shape,kw",
    [
        ("ns3d", (16, 12, 8), dict(nu_2=0.01, deltat0=0.02)),
        ("ns3d", (20, 15, 10), dict(nu_8=1e-4, deltat0=0.02, type_time_scheme="RK2", truncation_shape="spherical")),
        ("ns3d.strat", (16, 12, 8), dict(nu_4=0.001, deltat0=0.02, N=2.0, f=0.5)),
        ("ns2d", (32, 24), dict(nu_8=1e-6, deltat0=0.02, Lx=8, Ly=8)),
        ("ns2d", (32, 24), dict(nu_2=1e-3, deltat0=0.02, beta=0.3, type_time_scheme="RK2")),
        ("ns2d.strat", (32, 24), dict(nu_2=1e-3, deltat0=0.02, N=1.5, Lx=8, Ly=6)),
        ("ns2d.strat", (16, 32), dict(nu_8=1e-6, deltat0=0.02, N=0.7, type_time_scheme="RK2")),
        ("ns2d.bouss", (32, 24), dict(nu_4=1e-4, deltat0=0.02, Lx=8, Ly=8)),
    ],
)
def test_oracle_equals_reference_code(solver, shape, kw):
    """The restatement must reproduce the reference's own Python bit for bit."""
    nx, ny = shape[0], shape[1]
    nz = shape[2] if len(shape) == 3 else None
    ref = refshim.RefSim(solver, refshim.make_params(solver, nx, ny, nz, **kw))
    o = step_np.OracleSim(solver, nx, ny, nz, **kw)
    o.init_noise()
    if solver in ("ns2d.strat", "ns2d.bouss"):  # the noise recipe leaves b = 0: give it a field too
        o.set_state_spect(with_buoyancy_2d(o, solver, nx, ny, kw))
    ref.set_state_spect(np.array(o.state_spect))
    assert np.array_equal(ref.oper.where_dealiased, o.oper.where_dealiased)
    for _ in range(3):
        a = ref.step()
        b = np.array(o.one_time_step())
        assert np.array_equal(a, b)


@pytest.mark.skipif(not refshim.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize(
    "solver,shape,kw",
    [
        ("ns3d", (16, 12, 8), dict(type_time_scheme="Euler")),
        ("ns3d.strat", (16, 12, 8), dict(type_time_scheme="Euler_phaseshift", N=2.0)),
        ("ns2d", (32, 24), dict(type_time_scheme="Euler_phaseshift_random", Lx=8, Ly=8)),
        ("ns2d.strat", (16, 16), dict(type_time_scheme="Euler_phaseshift_random", nb_pairs=2, nb_steps_compute_new_pair=3, Lx=8, Ly=8)),
        ("ns3d", (16, 12, 8), dict(type_time_scheme="RK2_trapezoid")),
        ("ns3d.bouss", (16, 12, 8), dict(type_time_scheme="RK2_phaseshift")),
        ("ns2d", (32, 24), dict(type_time_scheme="RK2_phaseshift", beta=0.2, Lx=8, Ly=8)),
        ("ns3d", (16, 12, 8), dict(type_time_scheme="RK2_phaseshift_random")),
        ("ns3d.strat", (16, 12, 8), dict(type_time_scheme="RK2_phaseshift_random", N=2.0, nb_pairs=3, nb_steps_compute_new_pair=2)),
        ("ns3d.strat", (16, 12, 8), dict(type_time_scheme="RK2_phaseshift_exact", N=2.0)),
        ("ns2d.strat", (16, 16), dict(type_time_scheme="RK2_phaseshift_exact", N=1.5, Lx=8, Ly=8)),
    ],
)
def test_oracle_schemes_equal_reference_code(solver, shape, kw):
    """Euler / trapezoid / phase-shift schemes (pseudo_spect.py:245-796), weakly dealiased
    (coef_dealiasing = 0.9) so that the phase shifts -- and, for the random ones, the exact sequence of
    draws and pair renewals -- change the result; six steps, bit for bit."""
    import random

    nx, ny = shape[0], shape[1]
    nz = shape[2] if len(shape) == 3 else None
    kw = dict(nu_2=1e-2, deltat0=1e-2, coef_dealiasing=0.9, **kw)
    init = step_np.OracleSim(solver, nx, ny, nz, **{k: v for k, v in kw.items() if k != "type_time_scheme"})
    init.init_noise()
    s0 = np.array(init.state_spect)
    if solver in ("ns2d.strat", "ns2d.bouss"):
        s0 = with_buoyancy_2d(init, solver, nx, ny, {k: v for k, v in kw.items() if k != "type_time_scheme"})
    random.seed(12)  # both implementations draw from Python's `random`: run them one after the other
    ref = refshim.RefSim(solver, refshim.make_params(solver, nx, ny, nz, **kw))
    ref.set_state_spect(s0)
    ref_states = [ref.step() for _ in range(6)]
    random.seed(12)
    o = step_np.OracleSim(solver, nx, ny, nz, **kw)
    o.set_state_spect(s0)
    for a in ref_states:
        assert np.array_equal(a, np.array(o.one_time_step()))
    assert not np.array_equal(ref_states[0], s0)


@pytest.mark.skipif(not refshim.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("coef_group,forcing_rate", [(None, None), (1.0, None), (0.5, 8.0)])
def test_cfl_rule_ns2d_strat_equals_reference_code(coef_group, forcing_rate):
    """oracle CFL rule of ns2d.strat == the reference's TimeSteppingPseudoSpectralStrat
    (_init_compute_time_step + _compute_time_increment_CFL_uxuyb, ns2d/strat/time_stepping.py:31-186)."""
    refshim.install()
    from fluidsim.solvers.ns2d.strat.time_stepping import TimeSteppingPseudoSpectralStrat

    kw = dict(nu_2=1e-3, deltat0=0.2, N=3.0, Lx=8.0, Ly=6.0)
    params = refshim.make_params("ns2d.strat", 32, 24, None, **kw)
    params.time_stepping.USE_CFL = True
    params.time_stepping._set_attrib("cfl_coef_group", coef_group)
    if forcing_rate is not None:
        params.forcing.enable = True
        params.forcing._set_attrib("forcing_rate", forcing_rate)
    ref = refshim.RefSim("ns2d.strat", params)
    o = step_np.OracleSim("ns2d.strat", 32, 24, None, **kw)
    o.cfl_coef_group, o.forcing_rate = coef_group, forcing_rate
    o.init_noise(velo_max=0.05)
    o.set_state_spect(with_buoyancy_2d(o, "ns2d.strat", 32, 24, kw))
    ref.set_state_spect(np.array(o.state_spect))
    ts = TimeSteppingPseudoSpectralStrat.__new__(TimeSteppingPseudoSpectralStrat)
    ts.sim, ts.params = ref.sim, params
    ts._init_compute_time_step()
    lim = o.strat_time_increments()
    assert ts.deltat_dispersion_relation == lim["dispersion_relation"]
    if coef_group:
        assert ts.deltat_group_vel == lim["group_vel"] and ts.deltat_phase_vel == lim["phase_vel"]
    for scale in (1.0, 40.0, 41.0, 0.0):  # slow flow: a wave limit decides; fast flow: the advective CFL
        o.set_state_spect(scale * np.array(o.state_spect) if scale else 0 * np.array(o.state_spect))
        ref.set_state_spect(np.array(o.state_spect))
        ts.compute_time_increment_CLF()
        assert ts.deltat == o.compute_time_increment_CFL(cfl=ts.CFL, deltat_max=ts.deltat_max)


def test_spectrum3d_sums_to_energy():
    """Shell spectrum (restated fluidfft compute_3dspectrum): sum(E(k)) * deltak == energy."""
    o = step_np.OracleSim("ns3d", 16, 12, 10, nu_2=1e-2, Lx=5.0)
    o.init_noise()
    spec = o.compute_spectrum3d()
    assert abs(spec.sum() * o.oper.deltak - o.compute_energy()) < 1e-13 * o.compute_energy()


@pytest.mark.skipif(not refshim.available(), reason="/root/reference not present (GPU box)")
def test_cfl_rule_equals_reference_code():
    """oracle CFL time increment == the reference's own _compute_time_increment_CLF_uxuyuz /
    _compute_time_increment_CLF_from_tmp (base/time_stepping/base.py:320-354) on the same state."""
    refshim.install()
    from fluidsim.base.time_stepping.base import TimeSteppingBase

    o = step_np.OracleSim("ns3d", 16, 12, 8, nu_2=1e-2, deltat0=0.2)
    o.init_noise()
    ts = TimeSteppingBase.__new__(TimeSteppingBase)
    ts.CFL = 1.0
    ts.deltat_max = 0.2
    ts.deltat = 0.2
    phys = o.state_phys
    ts.sim = type("S", (), {})()
    ts.sim.oper = o.oper
    ts.sim.state = type("St", (), {"get_var": staticmethod(lambda key: phys.get_var(key))})()
    for _ in range(3):
        ts._compute_time_increment_CLF_uxuyuz()
        dt_o = o.compute_time_increment_CFL(cfl=1.0, deltat_max=0.2)
        assert ts.deltat == dt_o
        o.one_time_step()


def test_forced_goldens_differ_from_the_unforced_run():
    """The forcing really acts in the forced fixtures (same initial state as the unforced case)."""
    for forced, plain in (("ns3d_16x16x16_rk4_forced", "ns3d_16x16x16_rk4"), ("ns2d_32x32_rk4_forced", "ns2d_32x32_rk4")):
        _, zf = load_golden(forced)
        _, zp = load_golden(plain)
        assert np.array_equal(zf["state0"], zp["state0"])
        assert np.abs(zf["forcing"]).max() > 0
        assert rel_err(zf["state1"], zp["state1"]) > 1e-4


def test_oracle_spectra_and_means_are_consistent():
    """1-D / 3-D spectra [EXT restatement] integrate to the energy; means add up; the dissipation of a
    nu_2 run equals 2 nu_2 Z for a solenoidal field (Z = enstrophy)."""
    o = step_np.OracleSim("ns3d", 16, 12, 8, nu_2=1e-2, nu_m4=1e-3, Lx=6.0)
    o.init_noise()
    m = o.compute_spatial_means()
    sp = o.compute_spectra()
    op = o.oper
    assert abs(m["E"] - o.compute_energy()) < 1e-15
    for key, dk in (("E", op.deltak), ("E_kx", op.deltakx), ("E_ky", op.deltaky), ("E_kz", op.deltakz)):
        assert abs(sp[key].sum() * dk - m["E"]) < 1e-14 * m["E"]
    assert sp["vx_kx"].shape == (16 // 2 + 1,) and sp["vx_ky"].shape == (12 // 2 + 1,)
    assert sp["vx_kz"].shape == (8 // 2 + 1,)
    o2 = step_np.OracleSim("ns3d", 16, 16, 16, nu_2=1e-2)
    o2.init_noise()
    m2 = o2.compute_spatial_means()
    assert abs(m2["epsK"] - 2 * 1e-2 * o2.compute_enstrophy()) < 1e-12 * m2["epsK"]


# ------------------------------------------------------- invariants the reference's own tests assert
def _with_b(o, solver, shape, kw):
    """A state whose buoyancy field is non zero (the noise recipe of the 2-D solvers leaves b = 0)."""
    s = np.array(o.state_spect)
    if solver.startswith("ns2d"):
        s = with_buoyancy_2d(o, solver, shape[0], shape[1], kw)
    return s


def test_nonlinear_term_conserves_energy_with_buoyancy():
    """solvers/ns3d/strat/test_solver.py:26-58: sum Re(F_v . conj(v)) + Re(F_b conj(b)) / N^2 = 0."""
    o = step_np.OracleSim("ns3d.strat", 32, 16, 16, N=1.7)
    o.init_noise()
    # the noise recipe low-passes b but does not truncate it (ns3d/init_fields.py:124-142); on this small
    # grid the tail beyond the 2/3 cut-off is not negligible, and the identity needs a dealiased state
    o.dealiasing(o.state_spect)
    o.statephys_from_statespect()
    T, S = np.array(o.tendencies_nonlin()), np.array(o.state_spect)
    assert np.abs(S[3]).max() > 0
    T_tot = sum(np.real(T[i].conj() * S[i]) for i in range(3)) + np.real(T[3].conj() * S[3]) / o.N**2
    assert abs(o.oper.sum_wavenumbers(T_tot)) / o.oper.sum_wavenumbers(np.abs(T_tot)) < 1e-14


def test_ns2d_nonlinear_term_conserves_enstrophy_and_energy():
    """solvers/ns2d/test_solver.py:49-69: sum Re(conj(rot) F_rot) = 0 (enstrophy) and the same with the
    1 / K^2 weight (energy), inviscid, beta = 0."""
    o = step_np.OracleSim("ns2d", 48, 32, None, Lx=8.0, Ly=6.0)
    o.init_noise()
    rot = np.array(o.state_spect)[0]
    F = np.array(o.tendencies_nonlin())[0]
    sw = o.oper.sum_wavenumbers
    T = np.real(rot.conj() * F)
    assert abs(sw(T)) / sw(np.abs(T)) < 1e-14
    TE = T / o.oper.K2_not0
    assert abs(sw(TE)) / sw(np.abs(TE)) < 1e-14


@pytest.mark.parametrize("solver", ["ns2d.strat", "ns2d.bouss"])
def test_ns2d_buoyancy_solvers_conserve_energy(solver):
    """Simul.check_energy_conservation of ns2d.strat (solvers/ns2d/strat/solver.py:183-213):
    d/dt (E_K + E_A) = 0 for the inviscid tendencies, E_A = |b|^2 / (2 N^2).  ns2d.bouss has no
    background stratification: only the advective part of its terms is energy-neutral, checked as
    enstrophy-like invariance of b (sum Re(conj(b) F_b) = 0)."""
    kw = dict(N=1.3, Lx=8.0, Ly=6.0) if solver == "ns2d.strat" else dict(Lx=8.0, Ly=6.0)
    o = step_np.OracleSim(solver, 48, 32, None, **kw)
    o.init_noise()
    o.set_state_spect(_with_b(o, solver, (48, 32), kw))
    S, T = np.array(o.state_spect), np.array(o.tendencies_nonlin())
    sw = o.oper.sum_wavenumbers
    if solver == "ns2d.strat":
        division = 1.0 / o.oper.K2_not0
        division[0, 0] = 0
        pt = 0.5 * division * np.real(S[0].conj() * T[0]) + np.real(S[1].conj() * T[1]) / (2 * o.N**2)
    else:
        pt = np.real(S[1].conj() * T[1])
    assert abs(sw(pt)) / sw(np.abs(pt)) < 1e-14


@pytest.mark.parametrize(
    "scheme,order",
    [("Euler", 1), ("Euler_phaseshift", 1), ("RK2", 2), ("RK2_trapezoid", 2), ("RK2_phaseshift", 2),
     ("RK2_phaseshift_exact", 2), ("RK4", 4)],
)
def test_time_schemes_converge_at_their_order(scheme, order):
    """The counterpart of the reference's nl1d exact-solution test (solvers/nl1d/test_solver.py:56-163)
    on the hot-path solvers: halving the time step divides the error (against a fine RK4 run) by
    2^order."""
    kw = dict(nu_2=5e-2, Lx=2 * np.pi, Ly=2 * np.pi)
    init = step_np.OracleSim("ns2d", 32, 32, None, **kw)
    init.init_noise()
    s0 = np.array(init.state_spect)
    t_end = 0.4

    def run(sch, nsteps):
        o = step_np.OracleSim("ns2d", 32, 32, None, deltat0=t_end / nsteps, type_time_scheme=sch, **kw)
        o.set_state_spect(s0)
        for _ in range(nsteps):
            o.one_time_step()
        return np.array(o.state_spect)

    exact = run("RK4", 512)
    e1 = np.abs(run(scheme, 16) - exact).max()
    e2 = np.abs(run(scheme, 32) - exact).max()
    assert e1 > 1e-12  # resolvable error
    measured = np.log2(e1 / e2)
    assert order - 0.35 < measured < order + 0.6, (measured, e1, e2)


@pytest.mark.skipif(not refshim.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("name", all_golden_cases())
def test_committed_goldens_reproduce_from_the_reference_code(name):
    """Every committed fixture is what the reference's own modules produce today (same script, same
    inputs): guards against a stale or hand-edited golden."""
    import importlib.util
    import json

    spec = importlib.util.spec_from_file_location(
        "make_golden", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py")
    )
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    assert set(mg.CASES) == set(all_golden_cases())
    meta, arrays = mg.generate(name)
    stored_meta, z = load_golden(name)
    assert json.loads(json.dumps(meta)) == stored_meta
    for key, value in arrays.items():
        assert np.array_equal(np.asarray(value), z[key]), key
