"""Shared helpers for the parity tests."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def all_golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def golden_cases():
    """Unforced fixtures (the forced ones need a forcing object: ``forced_golden_cases``)."""
    return [n for n in all_golden_cases() if not n.endswith("_forced")]


def forced_golden_cases():
    return [n for n in all_golden_cases() if n.endswith("_forced")]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return meta, z


def rel_err(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def make_oracle(meta, stepping=False):
    """Numpy oracle for a golden / test case.  By default it only provides operators, tendencies and
    observables (built with RK4 when the case uses a *_random scheme, so that constructing it draws
    nothing from Python's `random`).  ``stepping=True``: the oracle of the case's own scheme, with the
    recorded ``random_seed`` applied right before construction, like the reference run that made the
    golden."""
    from oracle import step_np

    shape = meta["shape"]
    nz = shape[2] if len(shape) == 3 else None
    kw = dict(meta["params"])
    if stepping:
        if "random_seed" in meta:
            import random

            random.seed(meta["random_seed"])
    elif kw.get("type_time_scheme", "RK4").endswith("_random"):
        kw["type_time_scheme"] = "RK4"
    return step_np.OracleSim(meta["solver"], shape[0], shape[1], nz, **kw)


def make_gpu_sim(meta, fused=None, mask=None):
    """Build the GPU Simul for a golden / oracle case (same params names as the reference)."""
    import torch

    from fluidsim_b200.solvers import SIMUL_CLASSES

    solver = meta["solver"]
    cls = SIMUL_CLASSES[solver]
    p = cls.create_default_params()
    kw = dict(meta["params"])
    shape = meta["shape"]
    p.oper.nx, p.oper.ny = shape[0], shape[1]
    if len(shape) == 3:
        p.oper.nz = shape[2]
    for key in ("Lx", "Ly", "Lz", "coef_dealiasing", "truncation_shape"):
        if key in kw:
            setattr(p.oper, key, kw.pop(key))
    if not solver.startswith("ns2d"):
        p.oper.Lx = meta["params"].get("Lx", 2 * np.pi)
        p.oper.Ly = meta["params"].get("Ly", 2 * np.pi)
        p.oper.Lz = meta["params"].get("Lz", 2 * np.pi)
    else:
        p.oper.Lx = meta["params"].get("Lx", 2 * np.pi)
        p.oper.Ly = meta["params"].get("Ly", 2 * np.pi)
    p.time_stepping.USE_CFL = False
    p.time_stepping.type_time_scheme = kw.pop("type_time_scheme", "RK4")
    p.time_stepping.deltat0 = kw.pop("deltat0", 1e-2)
    for key in ("nb_pairs", "nb_steps_compute_new_pair"):  # pseudo_spect.py:159-167
        if key in kw:
            setattr(p.time_stepping.phaseshift_random, key, kw.pop(key))
    for key in ("nu_2", "nu_4", "nu_8", "nu_m4", "f", "N", "beta", "no_vz_kz0", "projection"):
        if key in kw:
            setattr(p, key, kw.pop(key))
    assert not kw, kw
    if "random_seed" in meta:  # *_random schemes draw from Python's `random` like the reference
        import random

        random.seed(meta["random_seed"])
    sim = cls(p, fused=fused)
    if mask is not None:
        # the dealiasing mask is an INPUT of the CUDA path (cubic comparator is [EXT] unpinned)
        sim.oper.where_dealiased = torch.from_numpy(np.ascontiguousarray(mask)).to(sim.oper.device)
    return sim


def set_state(sim, arr):
    import torch

    sim.state.state_spect.tensor.copy_(torch.from_numpy(np.ascontiguousarray(arr)))
    sim.state.mark_spect_modified()
