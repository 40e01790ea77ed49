"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against the
numpy oracle and the committed golden vectors made from the reference's own code.

Tolerances (BASELINE.json north_star): float64 state within 1e-10 relative per step; energy /
enstrophy / spectra within 1e-8 after 100 steps.  Observed differences are round-off (1e-15 ..
1e-13); the asserts below use 1e-11 per step so that regressions are caught early.
"""
import numpy as np
import pytest
import scipy.fft as sfft

from helpers import forced_golden_cases, golden_cases, load_golden, make_gpu_sim, make_oracle, rel_err, set_state

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-11  # north_star tolerance is 1e-10 relative per step
TOL_OBS = 1e-8


def _torch():
    import torch

    return torch


# ------------------------------------------------------------------ transforms
@pytest.mark.parametrize(
    "shape",
    [(8, 8, 8), (16, 32, 64), (128, 64, 32), (4, 11, 16), (10, 15, 20), (9, 8, 7), (16, 16, 6), (256, 256, 256)],
)
def test_fft3d_matches_scipy(shape):
    torch = _torch()
    from fluidsim_b200.fft import FFT3DWithB200

    o = FFT3DWithB200(*shape)
    rng = np.random.default_rng(1)
    x = rng.random(shape) - 0.5
    xd = torch.from_numpy(x).cuda()
    k = o.fft(xd)
    kr = sfft.rfftn(x) / np.prod(shape)
    assert rel_err(k.cpu().numpy(), kr) < 1e-14
    x2 = o.ifft(k)
    assert float((x2 - xd).abs().max()) < 1e-14
    # ifft_as_arg must leave its input intact; the destroy variant may not
    k0 = k.clone()
    out = o.create_arrayX()
    o.ifft_as_arg(k, out)
    assert torch.equal(k, k0)
    o.ifft_as_arg_destroy(k, out)
    assert float((out - xd).abs().max()) < 1e-14
    # host (numpy) calling convention of the reference
    kh = np.empty(o.get_shapeK_loc(), dtype=np.complex128)
    o.fft_as_arg(x, kh)
    assert rel_err(kh, kr) < 1e-14


@pytest.mark.parametrize("shape", [(8, 8), (256, 256), (64, 1024), (2048, 512), (16, 2048), (15, 24), (11, 9)])
def test_fft2d_matches_scipy(shape):
    torch = _torch()
    from fluidsim_b200.fft import FFT2DWithB200

    o = FFT2DWithB200(*shape)
    rng = np.random.default_rng(2)
    x = rng.random(shape) - 0.5
    xd = torch.from_numpy(x).cuda()
    k = o.fft(xd)
    assert rel_err(k.cpu().numpy(), sfft.rfft2(x) / np.prod(shape)) < 1e-14
    assert float((o.ifft(k) - xd).abs().max()) < 1e-14


def test_c2r_ignores_imaginary_part_of_self_conjugate_modes():
    """FFTW c2r semantics: Im of the kx=0 / Nyquist self-conjugate modes does not matter."""
    torch = _torch()
    from fluidsim_b200.fft import FFT3DWithB200

    shape = (8, 8, 16)
    o = FFT3DWithB200(*shape)
    rng = np.random.default_rng(3)
    k = rng.random(o.get_shapeK_loc()) + 1j * rng.random(o.get_shapeK_loc())
    ref = sfft.irfftn(k, s=shape) * np.prod(shape)
    out = o.ifft(torch.from_numpy(k).cuda()).cpu().numpy()
    assert rel_err(out, ref) < 1e-13


# ------------------------------------------------------------------ operators
def test_operators3d_against_oracle():
    torch = _torch()
    meta = dict(solver="ns3d", shape=(16, 12, 10), params=dict(Lx=6.0, Ly=5.0, Lz=3.0, nu_2=1e-2))
    o = make_oracle(meta)
    sim = make_gpu_sim(meta, fused=False)
    oper = sim.oper
    assert np.array_equal(oper.where_dealiased.cpu().numpy(), o.oper.where_dealiased)
    assert np.allclose(oper.K2.cpu().numpy(), o.oper.K2, rtol=1e-15)
    rng = np.random.default_rng(4)
    shapeK = o.oper.shapeK_loc
    v = [rng.random(shapeK) + 1j * rng.random(shapeK) for _ in range(3)]
    vd = [torch.from_numpy(a).cuda() for a in v]
    rot = oper.rotfft_from_vecfft(*vd)
    rot_o = o.oper.rotfft_from_vecfft(*v)
    for a, b in zip(rot, rot_o):
        assert rel_err(a.cpu().numpy(), b) < 1e-15
    assert rel_err(oper.divfft_from_vecfft(*vd).cpu().numpy(), o.oper.divfft_from_vecfft(*v)) < 1e-15
    vp = [a.copy() for a in v]
    o.oper.project_perpk3d(*vp)
    vpd = [a.clone() for a in vd]
    oper.project_perpk3d(*vpd)
    for a, b in zip(vpd, vp):
        assert rel_err(a.cpu().numpy(), b) < 1e-14
    # projection is idempotent and kills the divergence (test_operators3d.py:53-180)
    assert float(oper.divfft_from_vecfft(*vpd).abs().max()) < 1e-13
    # vector product writes into b
    shapeX = o.oper.shapeX_loc
    a3 = [rng.random(shapeX) for _ in range(3)]
    b3 = [rng.random(shapeX) for _ in range(3)]
    from fluidsim_b200.operators import vector_product

    ad = [torch.from_numpy(x).cuda() for x in a3]
    bd = [torch.from_numpy(x).cuda() for x in b3]
    r = vector_product(*ad, *bd)
    assert r[0] is bd[0]
    ref = np.cross(np.stack(a3, -1), np.stack(b3, -1))
    for i in range(3):
        assert rel_err(bd[i].cpu().numpy(), ref[..., i]) < 1e-15
    # energy reduction
    e = oper.compute_energy_from_K(vd[0])
    assert abs(e - o.oper.compute_energy_from_K(v[0])) < 1e-12 * abs(e)


def test_bad_inputs_raise():
    from fluidsim_b200.solvers import SimulNS3D

    p = SimulNS3D.create_default_params()
    p.oper.nx = p.oper.ny = p.oper.nz = 8
    p.time_stepping.type_time_scheme = "RK3"
    with pytest.raises(ValueError):
        SimulNS3D(p)
    p.time_stepping.type_time_scheme = "RK4"
    p.oper.type_fft = "fft3d.with_pyfftw"
    with pytest.raises(ValueError):
        SimulNS3D(p)
    p.oper.type_fft = "fft3d.with_b200"
    p.oper.truncation_shape = "banana"
    with pytest.raises(ValueError):
        SimulNS3D(p)


# ------------------------------------------------------------------ golden vectors (reference code)
@pytest.mark.parametrize("name", golden_cases())
@pytest.mark.parametrize("fused", [False, True])
def test_step_matches_reference_golden(name, fused):
    meta, z = load_golden(name)
    sim_probe = make_gpu_sim(meta, fused=False, mask=z["mask"])
    if fused and not sim_probe.oper.plan.is_fast:
        pytest.skip("fused path needs power-of-two sizes")
    if fused and not sim_probe.supports_fused:
        with pytest.raises(ValueError):  # ns2d.strat / ns2d.bouss: operator-level kernels only
            make_gpu_sim(meta, fused=True, mask=z["mask"])
        return
    sim = make_gpu_sim(meta, fused=fused, mask=z["mask"])
    set_state(sim, z["state0"])
    tend = sim.tendencies_nonlin_fused() if fused else sim.tendencies_nonlin()
    assert rel_err(tend.numpy(), z["tend0"]) < TOL_STEP
    sim.time_stepping.one_time_step()
    assert rel_err(sim.state.state_spect.numpy(), z["state1"]) < TOL_STEP
    for _ in range(meta["nsteps"] - 1):
        sim.time_stepping.one_time_step()
    assert rel_err(sim.state.state_spect.numpy(), z["stateN"]) < 10 * TOL_STEP
    if meta["solver"].startswith("ns2d"):
        e = sim.state.compute_energy_spect()  # sum' |rot|^2 / K2 / 2
        assert abs(e - float(z["energyN"])) < TOL_OBS * abs(float(z["energyN"]))
    else:
        e = sim.state.compute_energy_spect()
        assert abs(e - float(z["energyN"])) < TOL_OBS * abs(float(z["energyN"]))
        # physical-space energy (lazy state_phys = c2r of the state) against the oracle's.  Note:
        # E_phys == E_spect (base/state.py:385-392) only holds when the truncation removes the
        # Nyquist planes, which "spherical" on an anisotropic grid does not.
        o = make_oracle(meta)
        o.set_state_spect(z["stateN"])
        e_phys_o = 0.5 * float(np.mean(sum(np.array(o.state_phys[i]) ** 2 for i in range(3))))
        assert abs(sim.state.compute_energy_phys() - e_phys_o) < TOL_OBS * e_phys_o
        if meta["params"].get("truncation_shape", "cubic") == "cubic":
            assert sim.state.check_energy_equal_phys_spect()


# ------------------------------------------------------------------ BASELINE configs 1 and 2
def _run_both(meta, nsteps, init):
    o = make_oracle(meta)
    getattr(o, init)()
    sim = make_gpu_sim(meta, fused=True, mask=o.oper.where_dealiased)
    set_state(sim, np.array(o.state_spect))
    worst = 0.0
    for _ in range(nsteps):
        o.one_time_step()
        sim.time_stepping.one_time_step()
        worst = max(worst, rel_err(sim.state.state_spect.numpy(), np.array(o.state_spect)))
    return o, sim, worst


def test_config1_ns2d_256_noise_rk4_100_steps():
    """BASELINE config 1: ns2d 256^2 random init RK4 (fluidsim-bench protocol: L=8, nu_8=1)."""
    meta = dict(solver="ns2d", shape=(256, 256), params=dict(nu_8=1.0, deltat0=1e-3, Lx=8.0, Ly=8.0))
    o, sim, worst = _run_both(meta, 100, "init_noise")
    assert worst < 1e-10
    e_o, z_o = o.compute_energy(), o.compute_enstrophy()
    assert abs(sim.state.compute_energy_spect() - e_o) < TOL_OBS * e_o
    z_gpu = 0.5 * sim.oper.oper_fft.sum_wavenumbers_abs2(sim.state.state_spect.tensor)
    assert abs(z_gpu - z_o) < TOL_OBS * z_o


def _check_observables(sim, o, tol=TOL_OBS):
    """GPU one-pass reduction (b2_observables) against the oracle's restatement of the reference
    outputs: spatial means (E, Ex, Ey, Ez, epsK...), 3-D and 1-D spectra of every component."""
    sim._ensure_fused_buffers()  # pushes the viscosities
    obs = sim.oper.compute_observables(sim.state.state_spect.tensor[:3])
    means = o.compute_spatial_means()
    for key in ("E", "Ex", "Ey", "Ez", "epsK", "epsK_hypo", "epsK4", "epsK8"):
        assert abs(obs[key] - means[key]) <= tol * max(abs(means[key]), 1e-30) + 1e-300, key
    assert abs(obs["enstrophy"] - o.compute_enstrophy()) <= tol * o.compute_enstrophy()
    spec = o.compute_spectra()
    for comp in ("vx", "vy", "vz"):
        scale = max(spec["E"].max(), 1e-300)
        assert np.abs(obs[comp] - spec[comp]).max() <= tol * scale, comp
        for ax in ("kx", "ky", "kz"):
            ref = spec[f"{comp}_{ax}"]
            assert obs[f"{comp}_{ax}"].shape == ref.shape
            assert np.abs(obs[f"{comp}_{ax}"] - ref).max() <= tol * max(spec["E_" + ax].max(), 1e-300), (comp, ax)
    # spectra integrate to the energy
    oper = sim.oper
    assert abs(obs["E_spectrum3d"].sum() * oper.deltak - obs["E"]) < 1e-12 * obs["E"]
    assert abs(obs["E_kx"].sum() * oper.deltakx - obs["E"]) < 1e-12 * obs["E"]
    assert abs(obs["E_kz"].sum() * oper.deltakz - obs["E"]) < 1e-12 * obs["E"]


def test_config2_ns3d_128_taylor_green_rk4_100_steps():
    """BASELINE config 2: ns3d 128^3 Taylor-Green RK4 (nu_2 = 1/1600, dt = 1e-2), the full 100 steps of
    the north-star: state within 1e-10 per step, then energy, enstrophy, 3-D AND 1-D spectra and the
    spatial means within 1e-8 (the oracle needs ~1 s per step on the box's host cores)."""
    meta = dict(solver="ns3d", shape=(128, 128, 128), params=dict(nu_2=1 / 1600.0, deltat0=1e-2))
    nsteps = 100
    o, sim, worst = _run_both(meta, nsteps, "init_taylor_green")
    assert worst < 1e-10
    e_o = o.compute_energy()
    assert abs(sim.state.compute_energy_spect() - e_o) < TOL_OBS * e_o
    t = sim.state.state_spect.tensor
    e_fft = 0.5 * (t[0].abs() ** 2 + t[1].abs() ** 2 + t[2].abs() ** 2)
    spec_gpu = sim.oper.compute_3dspectrum(e_fft)
    spec_o = o.compute_spectrum3d()
    assert np.abs(spec_gpu - spec_o).max() < TOL_OBS * spec_o.max()
    _check_observables(sim, o)
    # Taylor-Green anchor: E(0) = 1/8 and the energy has decayed (doc/examples taylor-green)
    assert 0.11 < e_o < 0.125
    # spatial means of the physical fields (vx^2 ...), from the lazily computed state_phys
    phys = sim.state.state_phys.numpy()
    for i in range(3):
        m_o = float(np.mean(np.array(o.state_phys[i]) ** 2))
        assert abs(float(np.mean(phys[i] ** 2)) - m_o) < TOL_OBS * max(m_o, 1e-3)


@pytest.mark.parametrize("name", ["ns3d_16x12x8_rk4_odd", "ns3d_32x16x8_rk2_f", "ns3d_20x15x10_rk4_spherical"])
def test_observables_kernel_matches_oracle_small_and_odd_grids(name):
    """b2_observables on anisotropic / odd grids and non-unit boxes (bin folding, r2c weights with an
    odd nx, hypo-viscous dissipation) against the oracle after the golden's steps."""
    meta, z = load_golden(name)
    o = make_oracle(meta)
    o.set_state_spect(z["stateN"])
    sim = make_gpu_sim(meta, fused=None, mask=z["mask"])
    set_state(sim, z["stateN"])
    import ctypes as C

    from fluidsim_b200._lib import SOLVER_IDS, call, ptr

    call("b2_set_physics", sim.oper.plan.handle, SOLVER_IDS["ns3d"], *sim._physics_args(), ptr(sim.oper.where_dealiased))
    obs = sim.oper.compute_observables(sim.state.state_spect.tensor[:3])
    means = o.compute_spatial_means()
    for key in ("E", "Ex", "Ey", "Ez", "epsK", "epsK_hypo", "epsK4", "epsK8"):
        assert abs(obs[key] - means[key]) <= 1e-12 * max(abs(means[key]), 1e-30) + 1e-300, key
    spec = o.compute_spectra()
    for key in ("vx", "vy", "vz", "vx_kx", "vy_ky", "vz_kz", "vx_kz", "vz_ky"):
        assert np.abs(obs[key] - spec[key]).max() <= 1e-12 * spec["E"].max(), key


def test_ns3d_strat_128_noise_rk4():
    """ns3d.strat (config 4's solver) at 128^3 against the oracle: 10 RK4 steps with N = 1."""
    meta = dict(solver="ns3d.strat", shape=(128, 128, 128), params=dict(nu_8=1e-10, deltat0=5e-3, N=1.0))
    o, sim, worst = _run_both(meta, 10, "init_noise")
    assert worst < 1e-10
    assert abs(sim.state.compute_energy_spect() - o.compute_energy()) < TOL_OBS * o.compute_energy()
    _check_observables(sim, o)


# ------------------------------------------------------------------ invariants at larger sizes
def test_nonlinear_term_conserves_energy_256():
    """solvers/ns3d/test_solver.py:73-96 at a size the oracle would not finish quickly."""
    torch = _torch()
    meta = dict(solver="ns3d", shape=(256, 256, 256), params=dict(nu_2=0.0, deltat0=1e-3))
    sim = make_gpu_sim(meta, fused=True)
    g = torch.Generator(device="cuda").manual_seed(5)
    shapeX = sim.oper.shapeX_loc
    v = [sim.oper.fft(torch.rand(shapeX, dtype=torch.float64, device="cuda", generator=g) - 0.5) for _ in range(3)]
    sim.oper.project_perpk3d(*v)
    sim.oper.dealiasing(*v)
    sim.state.init_statespect_from(vx_fft=v[0], vy_fft=v[1], vz_fft=v[2])
    T = sim.tendencies_nonlin_fused().tensor
    S = sim.state.state_spect.tensor
    ratio = (T.conj() * S).real
    tot = sum(sim.oper.sum_wavenumbers(ratio[i]) for i in range(3))
    tot_abs = sum(sim.oper.sum_wavenumbers(ratio[i].abs()) for i in range(3))
    assert abs(tot) / tot_abs < 1e-13
    # tendencies are divergence free and dealiased
    assert float(sim.oper.divfft_from_vecfft(T[0], T[1], T[2]).abs().max()) < 1e-12
    assert float(T[:, sim.oper.where_dealiased.bool()].abs().max()) == 0.0


def test_cfl_time_increment_matches_oracle():
    meta = dict(solver="ns3d", shape=(32, 32, 32), params=dict(nu_2=1e-2, deltat0=0.2))
    o = make_oracle(meta)
    o.init_noise()
    sim = make_gpu_sim(meta, fused=True, mask=o.oper.where_dealiased)
    sim.params.time_stepping.USE_CFL = True
    sim.time_stepping.init_from_params()
    set_state(sim, np.array(o.state_spect))
    for _ in range(3):
        dt_o = o.compute_time_increment_CFL(cfl=1.0, deltat_max=0.2)
        o.one_time_step()
        sim.time_stepping.one_time_step()
        assert abs(sim.time_stepping.deltat - dt_o) < 1e-12 * dt_o
        assert rel_err(sim.state.state_spect.numpy(), np.array(o.state_spect)) < 1e-10


def test_nan_raises_value_error():
    torch = _torch()
    meta = dict(solver="ns3d", shape=(16, 16, 16), params=dict(nu_2=1e-2, deltat0=1e-2))
    sim = make_gpu_sim(meta, fused=True)
    sim.state.state_spect.tensor.fill_(0)
    sim.state.state_spect.tensor[0, 1, 1, 1] = float("nan")
    with pytest.raises(ValueError, match="nan at it"):
        sim.time_stepping.one_time_step()


# ------------------------------------------------------------------ dealias-pruned transforms
@pytest.mark.parametrize(
    "name",
    ["ns3d_16x16x16_rk4", "ns3d_32x16x8_rk2_f", "strat_16x16x16_rk4", "ns2d_32x32_rk4",
     "ns3d_16x16x16_rk4_spherical", "ns3d_32x16x16_rk2_nomultalias", "strat_16x16x8_rk4_spherical"],
)
def test_pruned_steps_match_golden_and_unpruned(name):
    """From the second step on the fused path skips everything outside the bounding box of the
    kept modes; results must be unchanged (the skipped data are exact zeros)."""
    meta, z = load_golden(name)
    sims = []
    for use_pruning in (True, False):
        sim = make_gpu_sim(meta, fused=True, mask=z["mask"])
        sim.use_pruning = use_pruning
        set_state(sim, z["state0"])
        for _ in range(meta["nsteps"]):
            sim.time_stepping.one_time_step()
        sims.append(sim)
    a, b = sims[0].state.state_spect.numpy(), sims[1].state.state_spect.numpy()
    assert rel_err(a, z["stateN"]) < 10 * TOL_STEP
    assert rel_err(a, b) < 1e-13
    # the pruned path was really used, and the dealiased region is exactly zero
    import ctypes as C

    from fluidsim_b200._lib import lib

    bounds = (C.c_int * 5)()
    lib.b2_get_pruning_bounds(sims[0].oper.plan.handle, bounds)
    assert bounds[4] < sims[0].oper.shapeK_loc[-1]
    assert np.abs(a[:, z["mask"].astype(bool)]).max() == 0.0


def test_pruned_128_taylor_green_and_noise():
    for init in ("init_taylor_green", "init_noise"):
        meta = dict(solver="ns3d", shape=(64, 128, 32), params=dict(nu_2=1e-3, deltat0=5e-3, Lx=3.0))
        o, sim, worst = _run_both(meta, 6, init)
        assert sim._state_dealiased
        assert worst < 1e-10


# ------------------------------------------------------------------ BASELINE sizes: properties
def test_512_properties_fft_roundtrip_energy_and_pruning():
    """At 512^3 (BASELINE config 3 size) the oracle is too slow; check size-independent properties:
    FFT round trip, Parseval (phys/spect energy), energy conservation of the nonlinear term,
    divergence-free dealiased state after steps, pruned == unpruned."""
    torch = _torch()
    meta = dict(solver="ns3d", shape=(512, 512, 512), params=dict(nu_8=1e-20, deltat0=1e-3))
    sim = make_gpu_sim(meta, fused=True)
    oper = sim.oper
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.rand(oper.shapeX_loc, dtype=torch.float64, device="cuda", generator=g) - 0.5
    k = oper.fft(x)
    assert float((oper.ifft(k) - x).abs().max()) < 5e-15
    # Parseval with the r2c-aware sum (sum_wavenumbers semantics)
    e_x = oper.compute_energy_from_X(x)
    e_k = oper.compute_energy_from_K(k)
    assert abs(e_x - e_k) < 1e-13 * e_x
    v = [k]
    for _ in range(2):
        x.uniform_(-0.5, 0.5, generator=g)
        v.append(oper.fft(x))
    del x
    oper.project_perpk3d(*v)
    oper.dealiasing(*v)
    sim.state.init_statespect_from(vx_fft=v[0], vy_fft=v[1], vz_fft=v[2])
    del v, k
    T = sim.tendencies_nonlin_fused().tensor
    S = sim.state.state_spect.tensor
    tot = tot_abs = 0.0
    for i in range(3):
        r = (T[i].conj() * S[i]).real
        tot += oper.sum_wavenumbers(r)
        tot_abs += oper.sum_wavenumbers(r.abs())
        del r
    assert abs(tot) / tot_abs < 1e-12
    del T
    s0 = S.clone()
    for _ in range(2):  # step 1 unpruned, step 2 pruned
        sim.time_stepping.one_time_step()
    a = S.clone()
    sim.use_pruning = False
    sim.state.state_spect.tensor.copy_(s0)
    sim.state.mark_spect_modified()
    del s0
    for _ in range(2):
        sim.time_stepping.one_time_step()
    assert float((a - S).abs().max()) <= 1e-13 * float(a.abs().max())
    assert float(oper.divfft_from_vecfft(S[0], S[1], S[2]).abs().max()) < 1e-10
    assert float(S[0][oper.where_dealiased.bool()].abs().max()) == 0.0


def test_longest_lines_2048_fused_against_oracle():
    """Line length 2048 (largest supported) through the fused path: ns2d with nx = 2048 (x pass) and
    ns2d with ny = 2048 (y pass), and a thin ns3d grid with nz = 2048 (z pass)."""
    for shape in ((2048, 16), (16, 2048)):
        meta = dict(solver="ns2d", shape=shape, params=dict(nu_2=1e-3, deltat0=1e-3, Lx=8.0, Ly=8.0))
        o, sim, worst = _run_both(meta, 3, "init_noise")
        assert worst < 1e-10, (shape, worst)
    meta = dict(solver="ns3d", shape=(16, 8, 2048), params=dict(nu_2=1e-3, deltat0=1e-3))
    o, sim, worst = _run_both(meta, 3, "init_noise")
    assert worst < 1e-10, worst


def test_pruning_is_not_used_on_an_undealiased_state():
    """A state with energy in a dealiased mode must go through the unpruned path (the reference
    evolves such a mode for one step before removing it)."""
    torch = _torch()
    meta, z = load_golden("ns3d_16x16x16_rk4")
    res = []
    for use_pruning in (True, False):
        sim = make_gpu_sim(meta, fused=True, mask=z["mask"])
        sim.use_pruning = use_pruning
        s0 = z["state0"].copy()
        idx = tuple(np.argwhere(z["mask"] == 1)[5])
        s0[(0,) + idx] = 0.3 + 0.1j
        set_state(sim, s0)
        sim.time_stepping.one_time_step()
        res.append(sim.state.state_spect.numpy())
    assert rel_err(res[0], res[1]) < 1e-14
    # and a clean state written from outside is recognised as dealiased: first step already pruned
    sim = make_gpu_sim(meta, fused=True, mask=z["mask"])
    set_state(sim, z["state0"])
    assert not sim._state_dealiased
    sim.time_stepping.one_time_step()
    assert rel_err(sim.state.state_spect.numpy(), z["state1"]) < TOL_STEP


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ns3d_16x16x16_rk4", "strat_16x16x16_rk4", "ns2d_32x32_rk4"])
def test_pruned_with_extra_fully_masked_rows(name):
    """Masks whose fully dealiased indices are NOT one contiguous run (NO_KY0 / NO_SHEAR_MODES zero
    the ky = 0 or kz = 0 rows on top of the Nyquist band, /root/reference/fluidsim/operators/
    operators3d.py:268-278, operators2d.py): only the run around n/2 may be pruned, the kept rows
    between the masked ones must still be transformed (kept_range)."""
    import ctypes as C

    from fluidsim_b200._lib import lib

    meta, z = load_golden(name)
    mask = z["mask"].copy()
    if mask.ndim == 3:
        mask[:, 0, :] = 1  # ky = 0 plane (NO_KY0-like)
        mask[0, :, :] = 1  # kz = 0 plane (NO_SHEAR_MODES-like)
        mask[3, :, :] = 1  # an isolated fully masked kz index
    else:
        mask[0, :] = 1     # ky = 0 row (NO_KY0)
        mask[5, :] = 1
    s0 = z["state0"].copy()
    s0[:, mask.astype(bool)] = 0.0
    res = []
    for use_pruning in (True, False):
        sim = make_gpu_sim(meta, fused=True, mask=mask)
        sim.use_pruning = use_pruning
        set_state(sim, s0)
        for _ in range(3):
            sim.time_stepping.one_time_step()
        res.append(sim.state.state_spect.numpy())
        if use_pruning:
            bounds = (C.c_int * 5)()
            lib.b2_get_pruning_bounds(sim.oper.plan.handle, bounds)
            n_ax0 = mask.shape[0] if mask.ndim == 3 else 1
            if mask.ndim == 3:
                assert 0 < bounds[0] < bounds[1] <= n_ax0       # band around n/2, not [0, ...)
            assert 0 < bounds[2] < bounds[3]
    a, b = res
    assert np.abs(b).max() > 0
    assert rel_err(a, b) < 1e-13
    assert np.abs(a[:, mask.astype(bool)]).max() == 0.0
    # the unfused operator-level path agrees too (independent of any pruning logic)
    sim = make_gpu_sim(meta, fused=False, mask=mask)
    set_state(sim, s0)
    for _ in range(3):
        sim.time_stepping.one_time_step()
    assert rel_err(a, sim.state.state_spect.numpy()) < 1e-11


@pytest.mark.gpu
def test_state_edits_through_the_reference_idioms_drop_pruning():
    """Editing state_spect through get_var / set_var / statespect_from_statephys (reference idioms,
    /root/reference/fluidsim/base/state.py:318-332) must invalidate the "state is dealiased"
    knowledge: the next step re-checks on the device and runs unpruned if needed."""
    meta, z = load_golden("ns3d_16x16x16_rk4")
    idx = tuple(np.argwhere(z["mask"] == 1)[7])

    def fresh():
        sim = make_gpu_sim(meta, fused=True, mask=z["mask"])
        set_state(sim, z["state0"])
        sim.time_stepping.one_time_step()
        assert sim._state_dealiased
        return sim

    # (a) in-place edit through a get_var view
    sim, ref = fresh(), fresh()
    ref.use_pruning = False
    for s in (sim, ref):
        v = s.state.state_spect.get_var("vx_fft")
        v[idx] = 0.25 - 0.5j
        assert not s._state_dealiased
        s.time_stepping.one_time_step()
    assert rel_err(sim.state.state_spect.numpy(), ref.state.state_spect.numpy()) < 1e-14
    # (b) edit of state_phys through a held reference + statespect_from_statephys
    sim, ref = fresh(), fresh()
    ref.use_pruning = False
    for s in (sim, ref):
        phys = s.state.state_phys
        phys.tensor[0] += 0.125 * phys.tensor[1] ** 2  # puts energy in dealiased modes
        s.state.statespect_from_statephys()
        assert not s._state_dealiased
        s.time_stepping.one_time_step()
    a, b = sim.state.state_spect.numpy(), ref.state.state_spect.numpy()
    assert rel_err(a, b) < 1e-14
    # (c) reading state_phys does not invalidate anything
    sim = fresh()
    _ = sim.state.state_phys
    _ = sim.state.compute_energy_spect()
    assert sim._state_dealiased


# ------------------------------------------------------------------ forcing
def _forced_sim(meta, z, fused):
    """GPU Simul with an in_script forcing that returns the golden's constant forcing_fft."""
    import torch

    from fluidsim_b200.solvers import SIMUL_CLASSES

    sim0 = make_gpu_sim(meta, fused=fused, mask=z["mask"])
    p = sim0.params
    p.forcing.enable = True
    p.forcing.type = "in_script"
    sim = SIMUL_CLASSES[meta["solver"]](p, fused=fused)
    sim.oper.where_dealiased = torch.from_numpy(np.ascontiguousarray(z["mask"])).to(sim.oper.device)
    keys = sim.state.keys_state_spect
    forcing = z["forcing"]

    def compute_forcing_fft_each_time(self):
        return {key: torch.from_numpy(np.ascontiguousarray(forcing[i])).to(sim.oper.device)
                for i, key in enumerate(keys) if np.abs(forcing[i]).max() > 0}

    sim.forcing.forcing_maker.monkeypatch_compute_forcing_fft_each_time(compute_forcing_fft_each_time)
    return sim


@pytest.mark.gpu
@pytest.mark.parametrize("name", forced_golden_cases())
@pytest.mark.parametrize("fused", [False, True])
def test_forced_step_matches_reference_golden(name, fused):
    """`tendencies_fft += forcing.get_forcing()` (/root/reference/fluidsim/solvers/ns3d/solver.py:243-244,
    strat/solver.py:210-211, ns2d/solver.py:190-191): golden made by the reference's own
    tendencies_nonlin with a stub forcing object (tests/golden/make_golden.py)."""
    meta, z = load_golden(name)
    sim = _forced_sim(meta, z, fused)
    assert sim.time_stepping.fused == fused
    set_state(sim, z["state0"])
    sim.forcing.compute()
    if fused:
        tend = sim.tendencies_nonlin_fused().numpy()
    else:
        tend = sim.tendencies_nonlin().numpy()
    assert rel_err(tend, z["tend0"]) < TOL_STEP
    sim.time_stepping.one_time_step()
    assert rel_err(sim.state.state_spect.numpy(), z["state1"]) < TOL_STEP
    for _ in range(meta["nsteps"] - 1):
        sim.time_stepping.one_time_step()
    assert rel_err(sim.state.state_spect.numpy(), z["stateN"]) < 10 * TOL_STEP


@pytest.mark.gpu
@pytest.mark.parametrize("ftype", ["tcrandom", "proportional"])
def test_normalised_forcing_injects_the_prescribed_rate(ftype):
    """Forced isotropic turbulence set-up of BASELINE config 3 (doc/examples/simul_ns3d_forced_isotropic.py)
    at 64^3: the normalised forcing must satisfy the injection identity of
    normalize_forcingc_2nd_degree_eq (/root/reference/fluidsim/base/forcing/specific.py:587-677),
    sum' Re(conj(v) f) + dt/2 sum' |f|^2 = forcing_rate, every step, and energy must grow accordingly."""
    from fluidsim_b200.solvers import SimulNS3D

    p = SimulNS3D.create_default_params()
    p.oper.nx = p.oper.ny = p.oper.nz = 64
    p.nu_2 = 1e-3
    p.time_stepping.USE_CFL = False
    p.time_stepping.deltat0 = 5e-3
    p.forcing.enable = True
    p.forcing.type = ftype
    p.forcing.nkmin_forcing = 3
    p.forcing.nkmax_forcing = 4
    p.forcing.forcing_rate = 0.5
    p.forcing.random_seed = 3
    sim = SimulNS3D(p, fused=True)
    o = make_oracle(dict(solver="ns3d", shape=(64, 64, 64), params=dict(nu_2=1e-3, deltat0=5e-3)))
    o.init_noise()
    set_state(sim, 0.3 * np.array(o.state_spect))
    oper = sim.oper
    e0 = sim.state.compute_energy_spect()
    dt = sim.time_stepping.deltat
    for it in range(4):
        sim.time_stepping.one_time_step()
        # identity on the forcing used for the step just done (state before the step is gone:
        # recompute with the current state through the maker, which is what the next step will use)
        sim.forcing._t_last_computed = -np.inf
        sim.forcing.compute()
        f = sim.forcing.get_forcing().tensor[:3]
        v = sim.state.state_spect.tensor[:3]
        p1 = sum(oper.sum_wavenumbers((f[i].conj() * v[i]).real) for i in range(3))
        p2 = dt / 2 * sum(oper.sum_wavenumbers(f[i].abs() ** 2) for i in range(3))
        if ftype == "tcrandom":
            assert abs(p1 + p2 - 0.5) < 1e-10
        else:  # proportional: Z ((1 + alpha dt)^2 - 1) / dt = P with Z the shell energy
            assert abs(p1 + p2 - 0.5) < 1e-10
        # the forcing is solenoidal and sits on kept low-wavenumber modes only
        div = oper.divfft_from_vecfft(f[0], f[1], f[2])
        assert float(div.abs().max()) < 1e-12
    assert sim.state.compute_energy_spect() > e0


@pytest.mark.gpu
def test_fluidfft_plugin_contract_numpy_operators_on_the_gpu_fft():
    """FFT-only drop-in: the numpy operator layer and the reference-shaped
    RK4 step of the oracle driven by ``FFT3DWithB200`` as their fluidfft plugin object, numpy arrays in
    and out (``fft_as_arg`` / ``ifft_as_arg`` / ``ifft_as_arg_destroy`` / shapes / k_adim /
    sum_wavenumbers) -- must equal the pure-numpy run."""
    from fluidsim_b200.fft3d_with_b200 import FFTclass
    from oracle import step_np

    nx, ny, nz = 32, 16, 8
    kw = dict(nu_2=1e-2, deltat0=1e-2, Lx=6.0, Ly=4.0, Lz=3.0)
    ref = step_np.OracleSim("ns3d", nx, ny, nz, **kw)
    ref.init_noise()
    o = step_np.OracleSim("ns3d", nx, ny, nz, **kw)
    plugin = FFTclass(nz, ny, nx)
    assert plugin.get_shapeK_seq() == ref.oper.shapeK_seq and plugin.get_dimX_K() == (0, 1, 2)
    mask = o.oper.where_dealiased.copy()
    o.oper.__init__(nx, ny, nz, kw["Lx"], kw["Ly"], kw["Lz"], fft=plugin, coef_dealiasing=2.0 / 3)
    o.oper.Lx, o.oper.Ly, o.oper.Lz = kw["Lx"], kw["Ly"], kw["Lz"]
    o.oper.where_dealiased = mask
    assert o.oper.type_fft == "fluidsim_b200.fft"
    o.set_state_spect(np.array(ref.state_spect))
    t_ref, t_gpu = np.array(ref.tendencies_nonlin()), np.array(o.tendencies_nonlin())
    assert rel_err(t_gpu, t_ref) < 1e-12
    for _ in range(3):
        ref.one_time_step()
        o.one_time_step()
    assert rel_err(np.array(o.state_spect), np.array(ref.state_spect)) < 1e-11
    assert abs(o.compute_energy() - ref.compute_energy()) < 1e-12 * ref.compute_energy()
