#!/usr/bin/env python
"""Benchmark of the pseudo-spectral RK4 step (BASELINE.json metric: ns3d RK4 steps/s and
grid-pts*steps/s), own arm (CUDA, libb200spectral) and reference arm (CPU, oracle port).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--size 512] [--solver ns3d]

Prints ONE JSON line (rank 0).  A "step" is one full RK4 time step (4 evaluations of the nonlinear
term = 36 3-D FFTs + epilogues) of the named solver on a synthetic noise field; protocol restated
from fluidsim-bench (/root/reference/fluidsim/util/console/util.py:147-215): L = 2 pi,
coef_dealiasing = 2/3, nu_8 = 1, deltat0 = 1e-4, USE_CFL = False, outputs off.
"""

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ns3d_rk4_grid_points_steps_per_s"
UNIT = "grid-pts*steps/s"

# algorithmic HBM bytes, in field passes F = 16*n0*n1*(n2/2+1) bytes
STEP_PASSES = {("ns3d", "RK4"): 195, ("ns3d.strat", "RK4"): 282, ("ns2d", "RK4"): 57, ("ns3d", "RK2"): 90}
CLASS_NAMES = ["first_inverse_pass", "y_inverse", "x_fused_c2r_cross_r2c", "y_forward", "z_forward",
               "rk_project_dealias_epilogue"]
SOLVER_COUNTS = {"ns3d": (3, 6, 3), "ns3d.strat": (4, 7, 6), "ns2d": (1, 4, 1)}  # nvar, n_in, n_out


def class_passes(solver, fx=1.0, fy=1.0, fz=1.0):
    """Algorithmic field passes (read + written, unit F) per LAUNCH of each kernel class of the fused
    path as built.  fx, fy, fz = kept fraction of kx columns / ky rows / kz
    rows (dealias-pruned transforms; 1 = unpruned).  ns2d has no z passes (fz = 1)."""
    nvar, nin, nout = SOLVER_COUNTS[solver]
    box = fx * fy * fz
    if solver == "ns2d":
        first = 1 * fx * fy + nin * fx          # y-inverse with the ns2d prologue: R rot (kept), W 4
        return [first, 0.0, nin * fx + nout * fx, nout * fx + nout * fx * fy, 0.0,
                ((2 + 2) + (3 + 2) + (3 + 2) + (2 + 1)) / 4.0 * box + box / 16.0]
    first = nin * box + nin * fx * fy            # z-inverse: R kept box, W all z of kept columns
    yinv = nin * fx * fy + nin * fx
    xp = nin * fx + nout * fx
    yfwd = nout * fx + nout * fx * fy
    zfwd = nout * fx * fy + nout * box
    extra = 2 if solver == "ns3d.strat" else 0   # b and vz of the stage input (buoyancy coupling)
    # RK epilogue, averaged over the 4 stages (+ the stage-0 curl kernel R3 W3): reads raw T (nout),
    # S, acc; writes acc, next stage input, its vorticity (3)
    # (5 launches per RK4 step in this class: the stage-0 curl kernel + 4 epilogues)
    rk = ((nout + nvar + extra) + (2 * nvar + 3) + 6
          + 2 * ((nout + 2 * nvar + extra) + (2 * nvar + 3))
          + ((nout + nvar + extra) + nvar)) / 5.0 * box + box / 16.0
    return [first, yinv, xp, yfwd, zfwd, rk]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", dest="n", type=int, default=1024, help="grid size per axis (power of two)")
    ap.add_argument("--solver", default="ns3d", choices=["ns3d", "ns3d.strat", "ns2d"])
    ap.add_argument("--scheme", default="RK4", choices=["RK4", "RK2"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-n", type=int, default=256, help="grid size of the bounded CPU sample")
    ap.add_argument("--allow-shrink", action="store_true",
                    help="halve the grid instead of failing when it does not fit device memory")
    ap.add_argument("--no-parity", action="store_true", help="skip the golden parity check before the timed region")
    return ap.parse_args()


def host_mem_available():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable"):
                    return int(line.split()[1]) * 1024
    except OSError:
        pass
    return 0


def e2e_pipelined(S, host_in, ts, sim, slab, nb, barrier):
    """Seconds per batch of `nb` independent batches streamed host -> device -> step -> host."""
    import torch

    host_out = torch.empty(S.shape, dtype=S.dtype, pin_memory=True)
    S_in, S_out = torch.empty_like(S), torch.empty_like(S)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for i in range(nb + 2):
        if i < nb:  # H2D of batch i
            with torch.cuda.stream(s_in):
                S_in.copy_(host_in, non_blocking=True)
        if 1 <= i <= nb:  # step of batch i-1 (S was loaded from S_in at the end of the previous round)
            if slab:
                sim.mark_spect_modified()
            else:
                sim.state.mark_spect_modified()
            ts.one_time_step()
        if i >= 2:  # D2H of the result of batch i-2
            with torch.cuda.stream(s_out):
                host_out.copy_(S_out, non_blocking=True)
        torch.cuda.synchronize()
        if 1 <= i <= nb:
            S_out.copy_(S)
        if i < nb:
            S.copy_(S_in)
    torch.cuda.synchronize()
    barrier()
    dt = (time.perf_counter() - t0) / nb
    del host_out, S_in, S_out
    torch.cuda.empty_cache()
    return dt


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(device_index)],
                stdout=self.tmp, stderr=subprocess.DEVNULL,
            )
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.tmp.name)
        except OSError:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------- CPU arm
def cpu_run(solver, scheme, n, nsteps, warmup, max_seconds=25.0):
    """Oracle port (reference Python restated + pocketfft, all host threads) on an n^3 sample."""
    from oracle import step_np

    cores = len(os.sched_getaffinity(0))
    kw = dict(nu_8=1.0, deltat0=1e-4, type_time_scheme=scheme)
    if solver == "ns2d":
        o = step_np.OracleSim(solver, n, n, None, Lx=8.0, Ly=8.0, **kw)
    else:
        o = step_np.OracleSim(solver, n, n, n, **kw)
    o.init_noise()
    for _ in range(warmup):
        o.one_time_step()
    t0 = time.perf_counter()
    done = 0
    while done < nsteps:
        o.one_time_step()
        done += 1
        if time.perf_counter() - t0 > max_seconds:
            break
    dt = (time.perf_counter() - t0) / done
    pts = n**o.ndim
    return dict(ms_per_step=dt * 1e3, steps=done, value=pts / dt, cores=cores, n=n)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_n
    r = cpu_run(args.solver, args.scheme, n, args.steps, args.warmup, max_seconds=200.0)
    sample = (f"{args.solver} {n}^{3 if args.solver != 'ns2d' else 2} {args.scheme} noise init, {r['steps']} steps "
              f"(bounded sample of the {args.n}^3 workload; value is per grid point so sizes compare)")
    line = {
        "impl": "reference",
        "metric": METRIC if args.solver == "ns3d" else f"{args.solver}_{args.scheme.lower()}_grid_points_steps_per_s",
        "value": r["value"],
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": r["steps"],
        "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"],
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{args.solver} {args.n}^3 {args.scheme} float64 (fluidsim-bench protocol)",
                   "cpu_sample_n": n},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": sample + "; reference-Python restated in numpy + scipy pocketfft "
                                            "(fluidfft/FFTW/MPI are not installable offline)"},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------- own arm
def make_sim(args, torch):
    from fluidsim_b200.solvers import SIMUL_CLASSES

    cls = SIMUL_CLASSES[args.solver]
    p = cls.create_default_params()
    n = args.n
    p.oper.nx = p.oper.ny = n
    if args.solver != "ns2d":
        p.oper.nz = n
    else:
        p.oper.Lx = p.oper.Ly = 8.0
    p.oper.coef_dealiasing = 2.0 / 3
    p.nu_8 = 1.0
    p.time_stepping.USE_CFL = False
    p.time_stepping.USE_T_END = False
    p.time_stepping.deltat0 = 1e-4
    p.time_stepping.type_time_scheme = args.scheme
    sim = cls(p, fused=True)
    sim.time_stepping.check_nan_period = 0  # checked once after the timed region
    init_noise_on_device(sim, torch)
    torch.cuda.empty_cache()  # the init temporaries (one physical field) go back before the step buffers come
    return sim


def init_noise_on_device(sim, torch, seed=42):
    """Noise recipe of solvers/ns3d/init_fields.py:153-196 with a device RNG (the NumPy recipe
    cannot allocate the large grids on the host): uniform noise -> project -> dealias -> low-pass
    tanh(k0 - K) -> rescale to velo_max = 1."""
    oper = sim.oper
    g = torch.Generator(device=oper.device).manual_seed(seed)
    nvar = sim.state.state_spect.nvar
    S = sim.state.state_spect.tensor
    x = oper.create_arrayX()
    for i in range(nvar):
        x.uniform_(-0.5, 0.5, generator=g)
        oper.fft_as_arg(x, S[i])
    del x
    S[(slice(None),) + (0,) * sim.ndim] = 0.0
    if sim.ndim == 3:
        oper.project_perpk3d(S[0], S[1], S[2])
    oper.dealiasing(sim.state.state_spect)
    k0 = 2 * 3.141592653589793 / (oper.Lx / 4.0)
    chunk = max(1, S.shape[1] // 16)
    kx2 = (oper._k2d if sim.ndim == 3 else oper._kxd) ** 2
    for a in range(0, S.shape[1], chunk):
        if sim.ndim == 3:
            K = torch.sqrt(kx2[None, None, :] + oper._k1d[None, :, None] ** 2 + oper._k0d[a:a + chunk, None, None] ** 2)
        else:
            K = torch.sqrt(kx2[None, :] + oper._kyd[a:a + chunk, None] ** 2)
        S[:, a:a + chunk] *= (1.0 + torch.tanh(2 * 3.141592653589793 * (k0 - K) / k0)) / 2.0
    # normalise with the spectral energy (max |v| would need 3 more physical fields at 1024^3)
    e = sim.state.compute_energy_spect() if sim.ndim == 3 else None
    if e:
        S *= (0.5 / e) ** 0.5 * 0.3
    sim.state.mark_spect_modified()


def make_slab_sim(args, torch, dist):
    """N > 1: slab decomposition (fluidsim_b200.slab.SlabSimul), same protocol, strong scaling."""
    import math

    from fluidsim_b200._lib import call, ptr, stream_ptr
    from fluidsim_b200.params import create_default_params
    from fluidsim_b200.slab import SlabSimul

    p = create_default_params(args.solver)
    n = args.n
    p.oper.nx = p.oper.ny = p.oper.nz = n
    p.nu_8 = 1.0
    p.time_stepping.USE_CFL = False
    p.time_stepping.deltat0 = 1e-4
    p.time_stepping.type_time_scheme = args.scheme
    sim = SlabSimul(args.solver, p, lean=True)
    dev = sim.device
    dk = 2 * math.pi / (2 * math.pi)
    kadim = lambda m: torch.fft.fftfreq(m, 1.0 / m, dtype=torch.float64, device=dev).abs()
    ky = dk * (kadim(n)[sim.rank::sim.world] if sim.cyclic else kadim(n)[sim.rank * sim.nyl:(sim.rank + 1) * sim.nyl])
    kz = dk * kadim(n)
    kx = dk * torch.arange(n // 2 + 1, dtype=torch.float64, device=dev)
    c = 2.0 / 3
    lim = c * dk * (n // 2 + 1)
    mask = ((ky >= lim)[:, None, None] | (kz >= lim)[None, :, None] | (kx >= lim)[None, None, :])
    sim.set_local_mask(mask.to(torch.uint8))
    g = torch.Generator(device=dev).manual_seed(42 + sim.rank)
    S = sim.state_spect
    k0 = 2 * math.pi / (2 * math.pi / 4.0)
    for i in range(sim.nvar):
        tmp = torch.view_as_real(S[i])
        tmp.uniform_(-0.5, 0.5, generator=g)
    for a in range(0, sim.nyl, max(1, sim.nyl // 8)):
        b = min(sim.nyl, a + max(1, sim.nyl // 8))
        K = torch.sqrt(ky[a:b, None, None] ** 2 + kz[None, :, None] ** 2 + kx[None, None, :] ** 2)
        S[:, a:b] *= (1.0 + torch.tanh(2 * math.pi * (k0 - K) / k0)) / 2.0
    if sim.rank == 0:
        S[:, 0, 0, 0] = 0.0
    call("b2_project_perpk3d", sim.handle, ptr(S[0]), ptr(S[1]), ptr(S[2]), stream_ptr())
    call("b2_dealias", sim.handle, ptr(S), sim.nvar, ptr(sim.where_dealiased), stream_ptr())
    e = sim.compute_energy()
    S *= (0.045 / e) ** 0.5
    return sim


def parity_check(args, torch, world, rank):
    """Step committed reference goldens (tests/golden/*.npz, produced by the reference's own code)
    through the SAME plan type the timed region uses (fused single-GPU plan, or the slab plan on all
    ranks) and return the max relative state error.  Puts parity evidence into the bench line of
    every N (the GPU test suite skips the multi-GPU cases on a single-GPU box)."""
    import numpy as np

    gdir = os.path.join(ROOT, "tests", "golden")
    names = {"ns3d": ["ns3d_16x16x16_rk4", "ns3d_32x16x8_rk2_f"], "ns3d.strat": ["strat_16x16x16_rk4"],
             "ns2d": ["ns2d_32x32_rk4"]}[args.solver]
    worst, cases = 0.0, []
    for name in names:
        path = os.path.join(gdir, name + ".npz")
        if not os.path.exists(path):
            continue
        z = np.load(path)
        meta = json.loads(str(z["meta"]))
        kw = dict(meta["params"])
        shape = meta["shape"]
        if world > 1:
            from fluidsim_b200.params import create_default_params
            from fluidsim_b200.slab import SlabSimul

            if shape[2] % world or shape[1] % world:
                continue
            p = create_default_params(meta["solver"])
            p.oper.nx, p.oper.ny, p.oper.nz = shape
        else:
            from fluidsim_b200.solvers import SIMUL_CLASSES

            p = SIMUL_CLASSES[meta["solver"]].create_default_params()
            p.oper.nx, p.oper.ny = shape[0], shape[1]
            if len(shape) == 3:
                p.oper.nz = shape[2]
        for key in ("Lx", "Ly", "Lz", "coef_dealiasing", "truncation_shape"):
            if key in kw:
                setattr(p.oper, key, kw.pop(key))
        p.time_stepping.USE_CFL = False
        p.time_stepping.type_time_scheme = kw.pop("type_time_scheme", "RK4")
        p.time_stepping.deltat0 = kw.pop("deltat0", 1e-2)
        for key in list(kw):
            setattr(p, key, kw.pop(key))
        if world > 1:
            # lean buffers need an exactly dealiased start: state1 (after one reference step); the
            # golden's state0 carries round-off in dealiased modes
            sim = SlabSimul(meta["solver"], p, lean=True)
            sim.set_mask_from_global(z["mask"])
            sim.set_state_from_global(z["state1"])
            for _ in range(meta["nsteps"] - 1):
                sim.one_time_step()
            got = sim.gather_state()
        else:
            sim = SIMUL_CLASSES[meta["solver"]](p, fused=True)
            sim.oper.where_dealiased = torch.from_numpy(np.ascontiguousarray(z["mask"])).to(sim.oper.device)
            sim.state.state_spect.tensor.copy_(torch.from_numpy(np.ascontiguousarray(z["state0"])))
            sim.state.mark_spect_modified()
            for _ in range(meta["nsteps"]):
                sim.time_stepping.one_time_step()
            got = sim.state.state_spect.numpy()
        ref = z["stateN"]
        err = float(np.abs(got - ref).max() / np.abs(ref).max())
        worst = max(worst, err)
        cases.append(f"{name} x{meta['nsteps']} steps")
        del sim
    torch.cuda.synchronize()
    return {"max_rel_err": worst, "cases": cases, "tolerance": 1e-10,
            "against": "reference goldens (tests/golden, generated by the reference's own modules)",
            "plan": "slab x%d (lean buffers)" % world if world > 1 else "single-GPU fused"}


def own_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (own arm) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    # development: B2_BENCH_FORCE_SLAB=1 runs the slab (multi-GPU) code path on ONE rank, which makes
    # its kernels profilable with ncu (never wrap a multi-rank command in ncu)
    force_slab = world == 1 and os.environ.get("B2_BENCH_FORCE_SLAB", "0") not in ("0", "")
    if world > 1 or force_slab:
        import datetime

        if force_slab:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            os.environ.setdefault("RANK", "0")
            os.environ.setdefault("WORLD_SIZE", "1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=180))
    from fluidsim_b200 import _lib

    hbm_peak, peak_src = peaks()
    # memory guard (single GPU): the fused path holds 12 (ns3d, aliased buffers) / 19 (strat) K fields
    # plus the mask; a grid that does not fit is an ERROR for the headline (no silent halving)
    size_note = None
    if args.solver != "ns2d" and world == 1 and not force_slab:  # (slab runs split the same grid)
        free_b, _total_b = torch.cuda.mem_get_info()
        nfields = {"ns3d": 12, "ns3d.strat": 19}[args.solver]
        while args.n > 64:
            need = (nfields + 0.2) * 16.0 * args.n * args.n * (args.n // 2 + 1)
            if need < free_b:
                break
            msg = (f"{args.solver} {args.n}^3 needs {need / 1e9:.0f} GB of device memory, {free_b / 1e9:.0f} GB free")
            if not args.allow_shrink:
                raise RuntimeError(msg + " (use --allow-shrink to halve the grid, or more GPUs)")
            size_note = msg + "; halved"
            args.n //= 2
    parity = None
    if not args.no_parity:
        parity = parity_check(args, torch, world, rank)
    slab = world > 1 or force_slab
    if slab:
        sim = make_slab_sim(args, torch, dist)
        ts = sim
        S = sim.state_spect
        ndim = 3
    else:
        sim = make_sim(args, torch)
        ts = sim.time_stepping
        S = sim.state.state_spect.tensor
        ndim = sim.ndim
    npts = float(args.n) ** ndim
    F = 16.0 * S[0].numel()  # local field pass (per GPU)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        ts.one_time_step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        ts.one_time_step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = npts / (ms_per_step * 1e-3)  # whole job: the N ranks advance ONE n^3 grid (strong scaling)
    if not bool(torch.isfinite(S.real.sum()).item()):
        raise RuntimeError("state became non finite during the benchmark")

    # kept fractions of the dealias-pruned transforms actually used in the timed steps
    kept = (1.0, 1.0, 1.0)
    if not slab and sim.use_pruning and sim._fused_mask is not None:
        import ctypes as C0

        bnd = (C0.c_int * 5)()
        _lib.lib.b2_get_pruning_bounds(sim.oper.plan.handle, bnd)
        n0, n1, nk = (1,) * (3 - ndim) + tuple(sim.oper.shapeK_loc)
        kept = (bnd[4] / nk, (bnd[2] + n1 - bnd[3]) / n1, (bnd[0] + n0 - bnd[1]) / n0)
    elif slab and sim.use_pruning and sim._prune is not None:
        keepx, kz_lo, kz_hi, _, _, gy_lo, gy_hi = sim._prune["args"]
        kept = (keepx / sim.nk, (gy_lo + sim.ny - gy_hi) / sim.ny, (kz_lo + sim.nz - kz_hi) / sim.nz)

    # ---- per-kernel-class timing (CUDA events on the launching stream inside the library)
    roofline = None
    classes = {}
    import ctypes as C

    # every rank runs the profiled steps (they contain collectives); rank 0 reports its own timings
    _lib.lib.b2_profile_reset()
    _lib.lib.b2_profile_enable(1)
    nprof = max(2, min(args.steps, 5))
    for _ in range(nprof):
        ts.one_time_step()
    torch.cuda.synchronize()
    if rank != 0:
        _lib.lib.b2_profile_enable(0)
    if rank == 0:
        msarr = (C.c_double * 6)()
        cnt = (C.c_longlong * 6)()
        _lib.lib.b2_profile_get(msarr, cnt, 6)
        _lib.lib.b2_profile_enable(0)
        passes = class_passes(args.solver, *kept)
        tot = sum(msarr)
        best = None
        nst = 4 if args.scheme == "RK4" else 2
        for i, name in enumerate(CLASS_NAMES):
            if cnt[i] == 0:
                continue
            # class_passes are per launch of the SINGLE-GPU structure (one launch per stage and class,
            # nst + 1 for the epilogue class with its stage-0 curl kernel); the slab path issues the
            # same work as several per-field / per-chunk launches, so rates are formed per STEP
            per_step_launches_1gpu = (nst + 1) if (i == 5 and args.solver != "ns2d") else nst
            bytes_step = passes[i] * F * per_step_launches_1gpu
            ms_step = msarr[i] / nprof
            ach = bytes_step / (ms_step * 1e-3) / 1e9
            classes[name] = {"launches_per_step": cnt[i] / nprof, "avg_ms": msarr[i] / cnt[i],
                             "ms_per_step": ms_step, "share": msarr[i] / tot,
                             "alg_bytes_per_step": bytes_step, "alg_GBps": ach, "frac": ach / hbm_peak}
            if best is None or msarr[i] > msarr[best]:
                best = i
        bname = CLASS_NAMES[best]
        # DRAM traffic of the dominant kernel: read from the committed summary of the `ncu --set full`
        # capture of this configuration (profiles/r2_ncu_traffic.json), never a literal
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")) as f:
                tj = json.load(f)
            key = f"{args.solver}:{args.n}:{world}:{bname}"
            if key in tj:
                traffic = tj[key]["dram_bytes_per_launch"]
        except (OSError, ValueError, KeyError):
            traffic = None
        lps = classes[bname]["launches_per_step"]
        roofline = {"bound": "hbm", "kernel": bname, "achieved": classes[bname]["alg_GBps"], "peak": hbm_peak,
                    "unit": "GB/s", "frac": classes[bname]["frac"], "traffic": traffic,
                    "peak_source": peak_src,
                    "alg_bytes_per_launch": classes[bname]["alg_bytes_per_step"] / lps,
                    "avg_launch_ms": classes[bname]["avg_ms"],
                    "note": "algorithmic bytes of the launch as built (dealias-pruned: kept kx fraction "
                            f"{kept[0]:.3f}, ky {kept[1]:.3f}, kz {kept[2]:.3f}); per-GPU figures (rank 0)"}
    barrier()

    # ---- end to end through the public API with HOST buffers (state in pinned host memory)
    e2e = None
    e2e_note = None
    if not args.no_e2e and host_mem_available() < 3 * S.numel() * 16 * (world if world > 1 else 1):
        e2e_note = "skipped: not enough host memory for a pinned copy of the state"
    elif not args.no_e2e:
        # every rank keeps ITS part of the state in pinned host memory and copies it in and out around
        # every step (the reference-facing call with host buffers); max over ranks
        host = torch.empty(S.shape, dtype=S.dtype, pin_memory=True)
        host.copy_(S)
        ksteps = max(2, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            S.copy_(host, non_blocking=True)
            if not slab:
                sim.state.mark_spect_modified()  # host data: the stepper re-checks that it is dealiased
            else:
                sim.mark_spect_modified()
            ts.one_time_step()
            host.copy_(S, non_blocking=True)
            torch.cuda.synchronize()
        barrier()
        dt = (time.perf_counter() - t0) / ksteps
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        nbytes = S.numel() * 16 * world
        e2e = {"value": npts / dt, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
               "ms_per_step": dt * 1e3, "steps": ksteps, "mode": "serial",
               "note": "whole-job bytes (all ranks); each rank copies its slab over its own PCIe link"}
        # The same work as a 3-stage pipeline over independent batches (a "step" = one pass of the hot
        # path over one batch): H2D of batch k+1 and D2H of the result of batch k-1 run on their own
        # streams (two copy engines, PCIe full duplex) under the step of batch k.  Needs two more state
        # buffers on the device.
        pipe = None
        try:
            torch.cuda.empty_cache()
            free_dev, _ = torch.cuda.mem_get_info()
            # single-GPU runs only: the multi-rank variant has not been exercised on a multi-GPU box
            ok = (world == 1 and free_dev > 2 * S.numel() * 16 + (6 << 30)
                  and host_mem_available() > 2 * S.numel() * 16)
            if ok:
                nbatch = max(6, min(args.steps, 12))
                pipe = e2e_pipelined(S, host, ts, sim, slab, nbatch, barrier)
        except Exception as exc:  # the serial figure above stands
            e2e["pipelined_note"] = f"pipelined run failed: {type(exc).__name__}: {str(exc)[:120]}"
            pipe = None
        if pipe is not None and 0 < pipe < dt:
            e2e.update({"value": npts / pipe, "ms_per_step": pipe * 1e3, "mode": "pipelined", "steps": nbatch,
                        "serial_ms_per_step": dt * 1e3, "serial_value": npts / dt, "serial_steps": ksteps,
                        "note": e2e["note"] + "; pipelined over independent batches: the H2D copy of batch "
                                "k+1 and the D2H copy of the result of batch k-1 overlap the step of batch k "
                                "(time = fill + batches + drain, divided by the number of batches)"})
        elif pipe is not None:
            e2e["pipelined_ms_per_step"] = pipe * 1e3
        del host

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_run(args.solver, args.scheme, args.cpu_n, 3, 1, max_seconds=25.0)
        cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                        "sample": f"{args.solver} {r['n']}^{ndim} {args.scheme}, {r['steps']} steps of the numpy/"
                                  f"pocketfft oracle port, {r['ms_per_step']:.0f} ms/step"}

    if rank == 0:
        nst = 4 if args.scheme == "RK4" else 2
        cp = class_passes(args.solver, *kept)
        built_passes = nst * sum(cp[:5]) + (5 if args.solver != "ns2d" else 4) * cp[5] if args.scheme == "RK4" else None
        as_built = None
        if built_passes:
            bb = built_passes * F
            as_built = {"kept_fraction_kx_ky_kz": kept, "field_passes_per_step": built_passes,
                        "alg_bytes_per_step_per_gpu": bb,
                        "achieved_GBps": bb / (ms_per_step * 1e-3) / 1e9,
                        "frac": bb / (ms_per_step * 1e-3) / 1e9 / hbm_peak}
        step_passes = STEP_PASSES.get((args.solver, args.scheme))
        step_alg = step_passes * F if step_passes else None  # per GPU (F is the local field pass)
        nfft = {"ns3d": 36, "ns3d.strat": 52}.get(args.solver, 0) if args.scheme == "RK4" else 0
        nvlink = None
        if world > 1:
            # bytes sent per GPU, per direction, per step: every 3-D FFT moves (P-1)/P of the local
            # field once; the pruned exchange carries only kept ky rows x kx < keepx columns
            nvl_full = nfft * F * (world - 1) / world
            nvl_bytes = nvl_full * kept[0] * kept[1]
            nvlink = {"bytes_per_gpu_per_step": nvl_bytes, "bytes_unpruned": nvl_full, "peak_GBps": 770.0,
                      "peak_source": "measured peer copy per direction (B200_PROFILING.md)",
                      "floor_ms": nvl_bytes / 770e9 * 1e3,
                      "frac_of_step": nvl_bytes / 770e9 * 1e3 / ms_per_step}
        line = {
            "metric": METRIC if args.solver == "ns3d" else f"{args.solver}_{args.scheme.lower()}_grid_points_steps_per_s",
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "steps_per_s": 1e3 / ms_per_step,
            "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": f"{args.solver} {args.n}^{ndim} {args.scheme} float64, noise init, nu_8=1, dt=1e-4, "
                            "USE_CFL=False (fluidsim-bench protocol; forcing off)",
                "n": args.n,
                "size_note": size_note,
                "state_bytes_per_gpu": S.numel() * 16,
                "l2_policy": "inputs larger than L2 (no flush needed)" if S.numel() * 16 > 200e6 else "L2-resident problem",
                "note": "timed step = full RK4 step incl. end-of-step projection + dealiasing; state_phys is lazy "
                        "on this path (the reference, and the CPU arm, refresh it with 3-4 inverse FFTs every "
                        "step) and the per-step NaN check is done once after the timed region",
                "parallelism": ("single GPU" if not force_slab else "slab code path on ONE rank (diagnostic)") if world == 1 else f"slab x{world} (z-slabs in X, ky-slabs in K; NCCL all-to-all; lean buffers)",
            },
            "nvlink": nvlink,
            "roofline": roofline,
            "step_roofline": {
                "model": "SURVEY.md section 8d fused 3-pass model (unpruned)",
                "model_field_passes": step_passes,
                "alg_bytes_per_step": step_alg,
                "achieved_GBps": step_alg / (ms_per_step * 1e-3) / 1e9 if step_alg else None,
                "frac": step_alg / (ms_per_step * 1e-3) / 1e9 / hbm_peak if step_alg else None,
            },
            "step_traffic_as_built": as_built,
            "kernel_classes": classes,
            "cpu_baseline": cpu_baseline,
            "e2e": e2e if e2e is not None else ({"note": e2e_note} if e2e_note else None),
            "gpu_launches": launches,
            "clocks": clocks,
            "parity_check": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1 or force_slab:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        own_arm(args)


if __name__ == "__main__":
    main()
