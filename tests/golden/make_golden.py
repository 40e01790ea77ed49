"""Generate golden vectors by running the REFERENCE's own hot-path code (build container only).

The reference tree keeps no stored vectors for this path, so these
fixtures are produced here by importing the unmodified reference modules from /root/reference
through ``oracle.refshim`` (only the absent fluidfft layer is the numpy restatement
``oracle.fluidfft_np``).  Run from the repo root:

    python tests/golden/make_golden.py

Each ``.npz`` holds: the case parameters (json), the initial ``state_spect``, the dealiasing mask
the reference used, ``state_spect`` after 1 and after ``nsteps`` reference steps, the reference's
``tendencies_nonlin`` of the initial state, and scalar observables.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import refshim, step_np  # noqa: E402

CASES = {
    # name: (solver, shape (nx, ny[, nz]), nsteps, params)
    "ns3d_16x16x16_rk4": ("ns3d", (16, 16, 16), 5, dict(nu_2=1e-2, deltat0=2e-2)),
    "ns3d_32x16x8_rk2_f": ("ns3d", (32, 16, 8), 5, dict(nu_8=1e-8, deltat0=1e-2, type_time_scheme="RK2", f=0.7, Lx=6.0, Ly=4.0, Lz=3.0)),
    "ns3d_16x12x8_rk4_odd": ("ns3d", (16, 12, 8), 3, dict(nu_4=2e-3, nu_8=1e-6, deltat0=2e-2, Lx=6.0)),
    "ns3d_20x15x10_rk4_spherical": ("ns3d", (20, 15, 10), 3, dict(nu_2=5e-3, nu_m4=1e-3, deltat0=2e-2, truncation_shape="spherical")),
    "strat_16x16x16_rk4": ("ns3d.strat", (16, 16, 16), 5, dict(nu_2=1e-2, deltat0=2e-2, N=2.0)),
    "strat_16x8x32_rk2": ("ns3d.strat", (16, 8, 32), 4, dict(nu_4=1e-3, deltat0=1e-2, N=0.5, f=0.3, type_time_scheme="RK2")),
    "ns3d_16x16x16_rk4_spherical": ("ns3d", (16, 16, 16), 4, dict(nu_2=5e-3, deltat0=2e-2, truncation_shape="spherical", coef_dealiasing=0.8)),
    "ns3d_32x16x16_rk2_nomultalias": ("ns3d", (32, 16, 16), 4, dict(nu_4=1e-4, deltat0=1e-2, type_time_scheme="RK2", truncation_shape="no_multiple_aliases", coef_dealiasing=0.9, Lx=8.0)),
    "strat_16x16x8_rk4_spherical": ("ns3d.strat", (16, 16, 8), 4, dict(nu_2=1e-2, deltat0=1e-2, N=1.5, truncation_shape="spherical")),
    "ns2d_32x32_rk4": ("ns2d", (32, 32), 5, dict(nu_8=1e-8, deltat0=2e-2, Lx=8.0, Ly=8.0)),
    "ns2d_64x32_rk2_beta": ("ns2d", (64, 32), 5, dict(nu_2=1e-3, deltat0=1e-2, beta=0.4, type_time_scheme="RK2", Lx=8.0, Ly=8.0)),
    "ns2d_24x15_rk4_odd": ("ns2d", (24, 15), 3, dict(nu_2=1e-3, deltat0=2e-2, Lx=8.0, Ly=8.0, truncation_shape="no_multiple_aliases")),
    # f-4 breadth: ns3d.bouss (solvers/ns3d/bouss/solver.py) and params.no_vz_kz0 (solver.py:260-263)
    "bouss_16x16x16_rk4": ("ns3d.bouss", (16, 16, 16), 4, dict(nu_2=1e-2, deltat0=2e-2)),
    "ns3d_16x16x16_rk4_novzkz0": ("ns3d", (16, 16, 16), 4, dict(nu_2=1e-2, deltat0=2e-2, no_vz_kz0=True)),
    "strat_16x16x8_rk2_novzkz0": ("ns3d.strat", (16, 16, 8), 4, dict(nu_4=1e-3, deltat0=1e-2, N=1.5, type_time_scheme="RK2", no_vz_kz0=True)),
    "ns3d_16x16x16_rk4_toroidal": ("ns3d", (16, 16, 16), 4, dict(nu_2=1e-2, deltat0=2e-2, projection="toroidal")),
    "ns3d_16x12x8_rk2_poloidal": ("ns3d", (16, 12, 8), 4, dict(nu_2=1e-2, deltat0=1e-2, projection="poloidal", type_time_scheme="RK2", Lx=6.0)),
    "strat_16x16x16_rk4_poloidal": ("ns3d.strat", (16, 16, 16), 3, dict(nu_2=1e-2, deltat0=2e-2, N=2.0, projection="poloidal")),
    # f-4: the other time schemes of TimeSteppingPseudoSpectral (pseudo_spect.py:245-796); the *_random
    # ones draw from Python's `random` (seeded with `random_seed` right before the objects are built).
    # The phase-shift cases use coef_dealiasing = 0.9: with the 2/3 rule the truncation already removes
    # every aliasing error and the result does not depend on the phase shifts at all.
    "ns3d_16x16x16_euler": ("ns3d", (16, 16, 16), 3, dict(nu_2=1e-2, deltat0=5e-3, type_time_scheme="Euler")),
    "ns3d_16x16x16_euler_phaseshift": ("ns3d", (16, 16, 16), 3, dict(coef_dealiasing=0.9, nu_2=1e-2, deltat0=5e-3, type_time_scheme="Euler_phaseshift")),
    "ns3d_16x12x8_rk2_trapezoid": ("ns3d", (16, 12, 8), 3, dict(nu_2=1e-2, deltat0=1e-2, type_time_scheme="RK2_trapezoid", Lx=6.0)),
    "ns3d_16x16x16_rk2_phaseshift": ("ns3d", (16, 16, 16), 3, dict(coef_dealiasing=0.9, nu_2=1e-2, deltat0=1e-2, type_time_scheme="RK2_phaseshift")),
    "strat_16x16x16_rk2_phaseshift_exact": ("ns3d.strat", (16, 16, 16), 3, dict(coef_dealiasing=0.9, nu_2=1e-2, deltat0=1e-2, N=2.0, type_time_scheme="RK2_phaseshift_exact")),
    "ns2d_32x32_rk2_phaseshift": ("ns2d", (32, 32), 4, dict(coef_dealiasing=0.9, nu_8=1e-8, deltat0=1e-2, Lx=8.0, Ly=8.0, type_time_scheme="RK2_phaseshift")),
    "ns3d_16x16x16_rk2_phaseshift_random": ("ns3d", (16, 16, 16), 5, dict(coef_dealiasing=0.9, nu_2=1e-2, deltat0=1e-2, type_time_scheme="RK2_phaseshift_random", random_seed=11)),
    "strat_16x16x16_rk2_phaseshift_random": ("ns3d.strat", (16, 16, 16), 4, dict(coef_dealiasing=0.9, nu_2=1e-2, deltat0=1e-2, N=2.0, type_time_scheme="RK2_phaseshift_random", random_seed=3)),
    "ns2d_32x32_euler_phaseshift_random": ("ns2d", (32, 32), 5, dict(coef_dealiasing=0.9, nu_8=1e-8, deltat0=5e-3, Lx=8.0, Ly=8.0, type_time_scheme="Euler_phaseshift_random", random_seed=5)),
    "ns2d_strat_32x32_euler_phaseshift_random_pairs2": ("ns2d.strat", (32, 32), 6, dict(coef_dealiasing=0.9, nu_2=1e-3, deltat0=5e-3, N=1.5, Lx=8.0, Ly=8.0, type_time_scheme="Euler_phaseshift_random", nb_pairs=2, nb_steps_compute_new_pair=3, random_seed=21)),
    # forced cases: a constant forcing_fft on the shell 2 <= |k|/dk <= 3.5 handed to the reference's
    # tendencies_nonlin through a stub `sim.forcing` (get_forcing()), forcing.enable = True
    "ns3d_16x16x16_rk4_forced": ("ns3d", (16, 16, 16), 5, dict(nu_2=1e-2, deltat0=2e-2)),
    "strat_16x16x16_rk4_forced": ("ns3d.strat", (16, 16, 16), 4, dict(nu_2=1e-2, deltat0=2e-2, N=2.0)),
    # f-4 breadth: ns2d.strat / ns2d.bouss (solvers/ns2d/strat/solver.py:71-181, bouss/solver.py:65-173);
    # the b field of the initial state is a second noise realisation (the noise recipe leaves b = 0)
    "ns2d_strat_32x32_rk4": ("ns2d.strat", (32, 32), 5, dict(nu_8=1e-8, deltat0=1e-2, N=1.5, Lx=8.0, Ly=8.0)),
    "ns2d_strat_32x16_rk2": ("ns2d.strat", (32, 16), 4, dict(nu_2=1e-3, deltat0=1e-2, N=0.7, type_time_scheme="RK2", Lx=8.0, Ly=5.0)),
    "ns2d_strat_24x15_rk4_odd": ("ns2d.strat", (24, 15), 3, dict(nu_2=1e-3, deltat0=2e-2, N=1.0, Lx=8.0, Ly=8.0)),
    "ns2d_bouss_32x32_rk4": ("ns2d.bouss", (32, 32), 5, dict(nu_4=1e-5, deltat0=1e-2, Lx=8.0, Ly=8.0)),
    "ns2d_32x32_rk4_forced": ("ns2d", (32, 32), 5, dict(nu_8=1e-8, deltat0=2e-2, Lx=8.0, Ly=8.0)),
}


def make_forcing(o, seed=7):
    """Deterministic forcing_fft: random complex amplitudes on the kept modes of the shell
    2 <= |k| / deltak <= 3.5 of the velocity components (ns2d: of rot_fft); zero elsewhere."""
    oper = o.oper
    rng = np.random.default_rng(seed)
    K = np.sqrt(oper.K2)
    dk = max(getattr(oper, "deltakx"), getattr(oper, "deltaky"), getattr(oper, "deltakz", 0.0))
    shell = (K >= 2 * dk) & (K <= 3.5 * dk) & (np.asarray(oper.where_dealiased) == 0)
    f = np.zeros(o.state_spect.shape, dtype=np.complex128)
    nforced = 1 if o.solver == "ns2d" else 3
    for v in range(nforced):
        amp = rng.standard_normal(K.shape) + 1j * rng.standard_normal(K.shape)
        f[v][shell] = 0.05 * amp[shell]
    return f


class _StubForcing:
    """What the reference's tendencies_nonlin needs from sim.forcing."""

    def __init__(self, forcing_fft):
        self.forcing_fft = forcing_fft

    def get_forcing(self):
        return self.forcing_fft

    def compute(self):
        pass


def generate(name):
    """Run the reference for one case; returns (meta dict, dict of arrays) -- what main() stores."""
    solver, shape, nsteps, kw = CASES[name]
    nx, ny = shape[0], shape[1]
    nz = shape[2] if len(shape) == 3 else None
    # initial condition: the reference's noise recipe (restated in step_np.init_noise, which
    # is itself checked against the reference in tests/test_oracle.py)
    kw = dict(kw)
    random_seed = kw.pop("random_seed", None)
    scheme = kw.get("type_time_scheme", "RK4")
    okw = dict(kw)
    if scheme not in ("RK2", "RK4"):
        okw["type_time_scheme"] = "RK4"  # the oracle only provides the initial state here
    o = step_np.OracleSim(solver, nx, ny, nz, **okw)
    o.init_noise()
    s0 = np.array(o.state_spect)
    if solver in ("ns2d.strat", "ns2d.bouss"):
        o2 = step_np.OracleSim(solver, nx, ny, nz, **okw)
        o2.init_noise(seed=7)
        s0[1] = 0.5 * np.array(o2.state_spect)[0]
    params = refshim.make_params(solver, nx, ny, nz, **kw)
    if random_seed is not None:
        import random

        random.seed(random_seed)
    ref = refshim.RefSim(solver, params)
    ref.set_state_spect(s0)
    extra = {}
    if name.endswith("_forced"):
        forcing = make_forcing(o)
        sov = type(ref.sim.state.state_spect)(like=ref.sim.state.state_spect, value=0.0)
        sov[...] = forcing
        ref.sim.is_forcing_enabled = True
        ref.sim.params.forcing.enable = True
        ref.sim.forcing = _StubForcing(sov)
        extra["forcing"] = forcing
    mask = np.array(ref.oper.where_dealiased)
    tend0 = np.array(ref.sim.tendencies_nonlin())
    states = []
    for _ in range(nsteps):
        states.append(ref.step())
    e = step_np.OracleSim(solver, nx, ny, nz, **okw)
    e.set_state_spect(states[-1])
    meta = dict(solver=solver, shape=shape, nsteps=nsteps, params=kw,
                **({} if random_seed is None else {"random_seed": random_seed}))
    arrays = dict(state0=s0, mask=mask, tend0=tend0, state1=states[0], stateN=states[-1],
                  energyN=e.compute_energy(), enstrophyN=e.compute_enstrophy(), **extra)
    return meta, arrays


def main():
    outdir = os.path.dirname(os.path.abspath(__file__))
    for name in CASES:
        if os.path.exists(os.path.join(outdir, name + ".npz")) and "--all" not in sys.argv:
            continue  # committed fixtures are kept byte-identical; --all regenerates everything
        meta, arrays = generate(name)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), meta=json.dumps(meta), **arrays)
        print(name, "ok", arrays["state0"].shape)


if __name__ == "__main__":
    main()
