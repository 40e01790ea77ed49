"""GPU branch of ``TimeSteppingPseudoSpectral`` (RK2 / RK4 with exact linear term).

Host-side mirror of ``/root/reference/fluidsim/base/time_stepping/base.py:96-354`` (main loop,
``one_time_step``, CFL time increment) and ``pseudo_spect.py:155-243,469-517,798-984`` (schemes).
Two execution modes, same results to round-off:

* **fused** (default when all sizes are powers of two in [8, 2048]): one call of the C ABI
  ``b2_time_step`` per step -- per stage: first inverse pass with the curl computed on load,
  y pass, fused x pass (c2r x6 -> v x omega -> r2c x3), forward y / z passes and ONE epilogue kernel
  doing buoyancy coupling + Leray projection + dealiasing + the exact-linear RK update.
* **unfused**: the reference's own sequence (``tendencies_nonlin`` + ``step_Euler`` /
  ``rk4_step1-3`` / ``step_like_RK2``) executed with the operator-level kernels; works for any
  grid size (odd, non power of two) and is what ``sim.tendencies_nonlin`` exposes.
"""

import signal
from time import time
from warnings import warn

import torch

from ._lib import SCHEME_IDS, call, ptr, stream_ptr
from .setofvariables import SetOfVariables


class ExactLinearCoefs:
    """pseudo_spect.py:103-152 -- ``exact = exp(-dt sigma)``, ``exact2 = exp(-dt/2 sigma)``."""

    def __init__(self, time_stepping):
        self.time_stepping = time_stepping
        sim = time_stepping.sim
        self.sim = sim
        oper = sim.oper
        self.exact = torch.empty(oper.shapeK_loc, dtype=torch.float64, device=oper.device)
        self.exact2 = torch.empty_like(self.exact)
        self.dt_old = None
        if sim.params.time_stepping.USE_CFL:
            self.get_updated_coefs = self.get_updated_coefs_CLF
        else:
            self.compute(time_stepping.deltat)
            self.get_updated_coefs = self.get_coefs

    def compute(self, dt):
        p = self.sim.params
        call(
            "b2_exact_coefs", self.sim.oper.plan.handle, p.nu_2, p.nu_4, p.nu_8, p.nu_m4, dt,
            ptr(self.exact), ptr(self.exact2), stream_ptr(),
        )
        self.dt_old = dt

    def get_updated_coefs_CLF(self):
        dt = self.time_stepping.deltat
        if self.dt_old != dt:
            self.compute(dt)
        return self.exact, self.exact2

    def get_coefs(self):
        return self.exact, self.exact2


class TimeSteppingPseudoSpectralB200:
    def __init__(self, sim, fused=None):
        self.params = sim.params
        self.sim = sim
        self.it = 0
        self.t = 0
        self._stop_signal_received = False
        self._has_to_stop = False
        self.max_elapsed = None
        try:
            signal.signal(signal.SIGUSR2, self._handler_signals)
        except (ValueError, AttributeError):
            warn("Cannot handle signals - is multithreading on?")
        can_fuse = sim.oper.plan.is_fast and getattr(sim, "supports_fused", True)
        self.fused = can_fuse if fused is None else bool(fused)
        if self.fused and not sim.oper.plan.is_fast:
            raise ValueError("fused time stepping needs power-of-two grid sizes in [8, 2048]")
        if self.fused and not can_fuse:
            raise ValueError(f"solver {sim.short_name} runs on the operator-level kernels only (fused=False)")
        self.init_from_params()

    def _handler_signals(self, signal_number, stack):
        print(f"signal {signal_number} received.")
        self._stop_signal_received = True

    # ---- init (pseudo_spect.py:173-224, base.py:246-304) --------------------------------------------
    def init_from_params(self):
        self._init_compute_time_step()
        self._init_time_scheme()
        self._exact_linear_coefs = None

    @property
    def exact_linear_coefs(self):
        if self._exact_linear_coefs is None:
            self._exact_linear_coefs = ExactLinearCoefs(self)
        return self._exact_linear_coefs

    def _init_compute_time_step(self):
        params_ts = self.params.time_stepping
        if params_ts.USE_CFL:
            if params_ts.cfl_coef is not None:
                self.CFL = params_ts.cfl_coef
            elif any(params_ts.type_time_scheme.startswith(s) for s in ["RK2", "Euler"]):
                self.CFL = 0.4
            elif params_ts.type_time_scheme.startswith("RK4"):
                self.CFL = 1.0
            else:
                raise ValueError("Problem name time_scheme")
        self.deltat = params_ts.deltat0
        self.deltat_max = params_ts.deltat_max
        self._maxbuf = torch.zeros(1, dtype=torch.float64, device=self.sim.oper.device)
        self._dt_dev = None
        self._vmax_dev = None
        self._flag_dev = None

    def _init_time_scheme(self):
        type_time_scheme = self.params.time_stepping.type_time_scheme
        # pseudo_spect.py:191-224.  RK2 / RK4 have the fused single-call path; the Euler, trapezoid and
        # phase-shifting schemes are compositions of N() and elementwise updates on the same kernels
        schemes = {
            "RK2": self._time_step_RK2,
            "RK4": self._time_step_RK4,
            "Euler": self._time_step_Euler,
            "Euler_phaseshift": self._time_step_Euler_phaseshift,
            "Euler_phaseshift_random": self._time_step_Euler_phaseshift_random,
            "RK2_trapezoid": self._time_step_RK2_trapezoid,
            "RK2_phaseshift": self._time_step_RK2_phaseshift,
            "RK2_phaseshift_random": self._time_step_RK2_phaseshift_random,
            "RK2_phaseshift_exact": self._time_step_RK2_phaseshift_exact,
        }
        if type_time_scheme not in schemes:
            raise ValueError(f'Problem name time_scheme ("{type_time_scheme}")')
        self._time_step_RK = schemes[type_time_scheme]
        self._scheme_id = SCHEME_IDS.get(type_time_scheme)
        if self._scheme_id is None:
            self._composed = self.fused  # N() through the fused kernels, the scheme composed here
            self.fused = False
        else:
            self._composed = False
        if type_time_scheme.endswith("_random"):
            self._init_phaseshift_random()
        self._phaseshift = None
        self._state_spect_tmp = None
        self._state_spect_tmp1 = None

    # ---- CFL (base.py:320-354) -----------------------------------------------------------------------
    def _max_abs(self, x):
        call("b2_max_abs", ptr(x), x.numel(), ptr(self._maxbuf), stream_ptr())
        return float(self._maxbuf.item())

    def compute_time_increment_CLF(self):
        get_var = self.sim.state.get_var
        oper = self.sim.oper
        if self.sim.ndim == 3:
            tmp = (
                self._max_abs(get_var("vx")) / oper.deltax
                + self._max_abs(get_var("vy")) / oper.deltay
                + self._max_abs(get_var("vz")) / oper.deltaz
            )
        else:
            tmp = self._max_abs(get_var("ux")) / oper.deltax + self._max_abs(get_var("uy")) / oper.deltay
        self._compute_time_increment_CLF_from_tmp(tmp)

    def _compute_time_increment_CLF_from_tmp(self, tmp):
        deltat_CFL = self.CFL / tmp if tmp > 0 else self.deltat_max
        maybe_new_dt = min(deltat_CFL, self.deltat_max)
        normalize_diff = abs(self.deltat - maybe_new_dt) / maybe_new_dt
        if normalize_diff > 0.02:
            self.deltat = maybe_new_dt

    # ---- main loop (base.py:144-244) ----------------------------------------------------------------
    def start(self):
        self.main_loop(print_begin=True, save_init_field=True)

    def main_loop(self, print_begin=False, save_init_field=False):
        params_stepping = self.params.time_stepping
        if self.max_elapsed is not None:
            self._time_should_stop = time() + self.max_elapsed
        if params_stepping.USE_T_END:
            while self.t < params_stepping.t_end and not self._has_to_stop:
                self.one_time_step()
        else:
            while self.it < params_stepping.it_end and not self._has_to_stop:
                self.one_time_step()

    def is_simul_completed(self):
        if self.params.time_stepping.USE_T_END:
            return self.t >= self.params.time_stepping.t_end
        return self.it >= self.params.time_stepping.it_end

    def one_time_step(self):
        if self.params.time_stepping.USE_CFL and not self.fused:
            self.compute_time_increment_CLF()  # fused path: decided on the device inside the step
        if self.sim.is_forcing_enabled:  # base.py:212-213
            self.sim.forcing.compute()
        if self.max_elapsed is not None and time() > self._time_should_stop:
            self._has_to_stop = True
        if self._stop_signal_received:
            self._has_to_stop = True
        self.one_time_step_computation()
        self.t += self.deltat
        self.it += 1

    # ---- one step -----------------------------------------------------------------------------------
    check_nan_period = 1

    def one_time_step_computation(self):
        """solvers/ns3d/time_stepping.py:8-20 (3-D) / pseudo_spect.py:236-243 (2-D)."""
        sim = self.sim
        state_spect = sim.state.state_spect
        if self.fused:
            sim._ensure_fused_buffers()
            if sim.use_pruning and not sim._state_dealiased and sim._fused_mask is not None:
                # state written from outside: one cheap pass tells whether it is dealiased already
                # (a few ms at 1024^3 against ~200 ms saved by the pruned transforms)
                if self._flag_dev is None:
                    self._flag_dev = torch.zeros(1, dtype=torch.int32, device=sim.oper.device)
                call("b2_check_dealiased", sim.oper.plan.handle, ptr(state_spect.tensor), state_spect.nvar,
                     ptr(sim._fused_mask), ptr(self._flag_dev), stream_ptr())
                sim._state_dealiased = int(self._flag_dev.item()) == 0
            prune = sim.use_pruning and sim._state_dealiased and sim._fused_mask is not None
            call("b2_set_pruning", sim.oper.plan.handle, 1 if prune else 0)
            if self.params.time_stepping.USE_CFL:
                # CFL on the device: max|v| from the stage-0 x pass -> deltat (2 % hysteresis) ->
                # RK epilogues; only the new deltat (one double) comes back to the host
                if self._dt_dev is None:
                    self._dt_dev = torch.full((1,), float(self.deltat), dtype=torch.float64, device=sim.oper.device)
                    self._vmax_dev = torch.zeros(3, dtype=torch.float64, device=sim.oper.device)
                else:
                    self._dt_dev.fill_(float(self.deltat))
                call(
                    "b2_time_step_cfl", sim.oper.plan.handle, self._scheme_id, float(self.CFL),
                    float(self.deltat_max), ptr(self._dt_dev), ptr(self._vmax_dev),
                    ptr(state_spect.tensor), stream_ptr(),
                )
                self.deltat = float(self._dt_dev.item())
            else:
                call(
                    "b2_time_step", sim.oper.plan.handle, self._scheme_id, float(self.deltat),
                    ptr(state_spect.tensor), stream_ptr(),
                )
        else:
            self._time_step_RK()
            if sim.ndim == 3:
                sim.project_state_spect(state_spect)
            sim.oper.dealiasing(state_spect)
        sim.state.statephys_from_statespect()
        sim._state_dealiased = True  # every step ends with oper.dealiasing(state_spect)
        if self.check_nan_period and (self.it + 1) % self.check_nan_period == 0:
            call("b2_sum", ptr(state_spect.tensor), 2 * state_spect.tensor[0].numel(), ptr(self._maxbuf), stream_ptr())
            if torch.isnan(self._maxbuf).item():
                raise ValueError(f"nan at it = {self.it}, t = {self.t:.4f}")

    # ---- unfused schemes (pseudo_spect.py:469-517, 798-984) -------------------------------------------
    def _tmp_like_state(self, name):
        buf = getattr(self, name)
        if buf is None:
            buf = SetOfVariables(like=self.sim.state.state_spect)
            setattr(self, name, buf)
        return buf

    # ---- Euler / trapezoid / phase-shifting schemes (pseudo_spect.py:245-468, 519-796) -----------------
    def _compute_tendencies(self, state_spect=None, old=None):
        """N(state) -- through the fused kernels when the grid allows (arbitrary input: unpruned)."""
        sim = self.sim
        if self._composed:
            return sim.tendencies_nonlin_fused(state_spect, old=old)
        return sim.tendencies_nonlin(state_spect, old=old)

    def _like_state(self, tensor):
        return SetOfVariables(input_array=tensor, keys=self.sim.state.state_spect.keys, info="tmp")

    def _get_phaseshift(self):
        """pseudo_spect.py:281-300: exp(i/2 (dx Kx + dy Ky [+ dz Kz]))."""
        if self._phaseshift is None:
            oper = self.sim.oper
            if self.sim.ndim == 2:
                phase = 0.5 * (oper.deltax * oper.KX + oper.deltay * oper.KY)
            else:
                phase = 0.5 * (oper.deltax * oper.Kx + oper.deltay * oper.Ky + oper.deltaz * oper.Kz)
            self._phaseshift = torch.exp(1j * phase)
        return self._phaseshift

    def _init_phaseshift_random(self):
        """pseudo_spect.py:302-328."""
        pp = self.params.time_stepping.phaseshift_random
        if pp.nb_steps_compute_new_pair is None:
            pp.nb_steps_compute_new_pair = 2 if pp.nb_pairs == 1 else 4 * pp.nb_pairs
        self._index_phaseshift = 1
        self._previous_index_pair = 0
        self._previous_index_flip = 0
        self._pairs_phaseshift = [self._new_random_pair() for _ in range(pp.nb_pairs)]

    def _new_random_pair(self):
        """oper.get_phases_random (operators3d.py:1128-1146, operators2d) + compute_phaseshift_terms."""
        from random import uniform

        oper = self.sim.oper
        nd = self.sim.ndim
        alphas = tuple(uniform(-0.5, 0.5) for _ in range(nd))
        betas = tuple(a + 0.5 if a < 0 else a - 0.5 for a in alphas)
        if nd == 3:
            grids = (oper.deltax * oper.Kx, oper.deltay * oper.Ky, oper.deltaz * oper.Kz)
        else:
            grids = (oper.deltax * oper.KX, oper.deltay * oper.KY)
        phase_alpha = sum(a * g for a, g in zip(alphas, grids))
        phase_beta = sum(b * g for b, g in zip(betas, grids))
        return torch.exp(1j * phase_alpha), torch.exp(1j * phase_beta)

    def _get_phaseshift_random(self):
        """pseudo_spect.py:330-372."""
        from random import randint

        pp = self.params.time_stepping.phaseshift_random
        nb_pairs, nb_steps = pp.nb_pairs, pp.nb_steps_compute_new_pair
        if nb_pairs == 1 and nb_steps == 1:
            alpha, beta = self._pairs_phaseshift[0]
        elif nb_pairs == 1 and nb_steps == 2:
            pair = self._pairs_phaseshift[0]
            alpha, beta = pair if self._index_phaseshift == 1 else pair[::-1]
        else:
            index_pair = randint(0, nb_pairs - 1)
            pair = self._pairs_phaseshift[index_pair]
            index_flip = randint(0, 1)
            if index_pair == self._previous_index_pair and index_flip == self._previous_index_flip:
                index_flip = 0 if index_flip else 1
            self._previous_index_pair = index_pair
            self._previous_index_flip = index_flip
            alpha, beta = pair if index_flip else pair[::-1]
        if self._index_phaseshift == nb_steps:
            # the reference re-binds (alpha, beta) to the arrays of the oldest pair and overwrites them
            # in place with the new phases (:362-369): this step already uses the NEW pair
            self._index_phaseshift = 1
            self._pairs_phaseshift.pop(0)
            alpha, beta = self._new_random_pair()
            self._pairs_phaseshift.append((alpha, beta))
        else:
            self._index_phaseshift += 1
        return alpha, beta

    def _shifted_tendencies(self, phaseshift, state_tensor):
        """N(phaseshift * S) / phaseshift as a tensor."""
        shifted = self._like_state(phaseshift * state_tensor)
        return self._compute_tendencies(shifted).tensor / phaseshift

    def _euler_inplace(self, tendencies_tensor, diss):
        """step_Euler_inplace (pseudo_spect.py:59-61)."""
        S = self.sim.state.state_spect
        t = tendencies_tensor.contiguous()
        call("b2_step_euler", self.sim.oper.plan.handle, ptr(S.tensor), self.deltat, ptr(t), ptr(diss),
             ptr(S.tensor), S.nvar, stream_ptr())

    def _time_step_Euler(self):
        diss = self.exact_linear_coefs.get_updated_coefs()[0]
        self._euler_inplace(self._compute_tendencies().tensor, diss)

    def _time_step_Euler_phaseshift(self):
        diss = self.exact_linear_coefs.get_updated_coefs()[0]
        S = self.sim.state.state_spect.tensor
        tendencies_0 = self._compute_tendencies().tensor
        tendencies_shifted = self._shifted_tendencies(self._get_phaseshift(), S)
        self._euler_inplace(0.5 * (tendencies_0 + tendencies_shifted), diss)

    def _time_step_Euler_phaseshift_random(self):
        diss = self.exact_linear_coefs.get_updated_coefs()[0]
        S = self.sim.state.state_spect.tensor
        alpha, beta = self._get_phaseshift_random()
        t_alpha = self._shifted_tendencies(alpha, S)
        t_beta = self._shifted_tendencies(beta, S)
        self._euler_inplace(0.5 * (t_alpha + t_beta), diss)

    def _rk2_first_half(self, tendencies_0, diss):
        """state_spect_1 = step_Euler(state_spect, dt, tendencies_0, diss)."""
        sim = self.sim
        S = sim.state.state_spect
        state_spect_1 = self._tmp_like_state("_state_spect_tmp")
        call("b2_step_euler", sim.oper.plan.handle, ptr(S.tensor), self.deltat, ptr(tendencies_0.contiguous()),
             ptr(diss), ptr(state_spect_1.tensor), S.nvar, stream_ptr())
        return state_spect_1

    def _step_like_rk2(self, tendencies_d, diss, diss2):
        S = self.sim.state.state_spect
        call("b2_step_like_rk2", self.sim.oper.plan.handle, ptr(S.tensor), self.deltat,
             ptr(tendencies_d.contiguous()), ptr(diss), ptr(diss2), S.nvar, stream_ptr())

    def _time_step_RK2_trapezoid(self):
        dt = self.deltat
        diss, diss2 = self.exact_linear_coefs.get_updated_coefs()
        S = self.sim.state.state_spect.tensor
        tendencies_0 = self._compute_tendencies().tensor.clone()
        state_spect_1 = self._rk2_first_half(tendencies_0, diss)
        tendencies_1 = self._compute_tendencies(state_spect_1).tensor
        S.copy_((S + dt / 2 * tendencies_0) * diss + dt / 2 * tendencies_1)

    def _time_step_RK2_phaseshift(self):
        diss, diss2 = self.exact_linear_coefs.get_updated_coefs()
        tendencies_0 = self._compute_tendencies().tensor.clone()
        state_spect_1 = self._rk2_first_half(tendencies_0, diss)
        phaseshift = self._get_phaseshift()
        tendencies_1_shift = self._compute_tendencies(self._like_state(phaseshift * state_spect_1.tensor)).tensor
        tendencies_d = 0.5 * (tendencies_0 + tendencies_1_shift / phaseshift)
        self._step_like_rk2(tendencies_d, diss, diss2)

    def _time_step_RK2_phaseshift_random(self):
        """pseudo_spect.py:640-733."""
        diss, diss2 = self.exact_linear_coefs.get_updated_coefs()
        S = self.sim.state.state_spect.tensor
        alpha, beta = self._get_phaseshift_random()
        # both evaluations write their result over their input in the reference (:687-702)
        tendencies_0 = self._shifted_tendencies_aliased(alpha, S) / alpha
        state_spect_1 = self._rk2_first_half(tendencies_0, diss)
        tendencies_1 = self._shifted_tendencies_aliased(beta, state_spect_1.tensor) / beta
        tendencies_d = 0.5 * (tendencies_0 + tendencies_1)
        self._step_like_rk2(tendencies_d, diss, diss2)

    def _shifted_tendencies_aliased(self, phaseshift, state_tensor):
        """The reference evaluates ``compute_tendencies(state_spect_shift, old=state_spect_shift)``
        (pseudo_spect.py:764-766, 777-779): the output array IS the input array.  For ns3d.strat this
        changes the reference's result -- ``fb_fft`` is formed from ``vz_fft`` after ``fz_fft + b_fft``
        was written over it (strat/solver.py:198-207) -- so that solver evaluates the aliased call through
        the same operator-level sequence to stay identical to the reference; the other solvers read
        everything they need before the first write and may use the fused kernels."""
        sim = self.sim
        shifted = self._like_state(phaseshift * state_tensor)
        if sim.short_name == "ns3d.strat":
            return sim.tendencies_nonlin(shifted, old=shifted).tensor
        return self._compute_tendencies(shifted).tensor

    def _time_step_RK2_phaseshift_exact(self):
        """pseudo_spect.py:735-796."""
        diss, diss2 = self.exact_linear_coefs.get_updated_coefs()
        S = self.sim.state.state_spect.tensor
        phaseshift = self._get_phaseshift()
        tendencies_0 = self._compute_tendencies(self._like_state(S)).tensor
        tendencies_0_shift = self._shifted_tendencies_aliased(phaseshift, S)
        tendencies_d0 = 0.5 * (tendencies_0 + tendencies_0_shift / phaseshift)
        state_spect_1 = self._rk2_first_half(tendencies_d0, diss)
        tendencies_1 = self._compute_tendencies(state_spect_1).tensor
        tendencies_1_shift = self._shifted_tendencies_aliased(phaseshift, state_spect_1.tensor)
        tendencies_d = 0.5 * (tendencies_d0 + 0.5 * (tendencies_1 + tendencies_1_shift / phaseshift))
        self._step_like_rk2(tendencies_d, diss, diss2)

    def _time_step_RK2(self):
        dt = self.deltat
        diss, diss2 = self.exact_linear_coefs.get_updated_coefs()
        sim = self.sim
        h = sim.oper.plan.handle
        state_spect = sim.state.state_spect
        nvar = state_spect.nvar
        tendencies_0 = sim.tendencies_nonlin()
        state_spect_12 = self._tmp_like_state("_state_spect_tmp")
        call("b2_step_euler", h, ptr(state_spect.tensor), dt / 2, ptr(tendencies_0.tensor), ptr(diss2),
             ptr(state_spect_12.tensor), nvar, stream_ptr())
        tendencies_12 = sim.tendencies_nonlin(state_spect_12, old=tendencies_0)
        call("b2_step_like_rk2", h, ptr(state_spect.tensor), dt, ptr(tendencies_12.tensor), ptr(diss),
             ptr(diss2), nvar, stream_ptr())

    def _time_step_RK4(self):
        dt = self.deltat
        diss, diss2 = self.exact_linear_coefs.get_updated_coefs()
        sim = self.sim
        h = sim.oper.plan.handle
        state_spect = sim.state.state_spect
        nvar = state_spect.nvar
        S = ptr(state_spect.tensor)
        tendencies_0 = sim.tendencies_nonlin()
        state_spect_tmp = self._tmp_like_state("_state_spect_tmp")
        state_spect_tmp1 = self._tmp_like_state("_state_spect_tmp1")
        acc, tmp1 = ptr(state_spect_tmp.tensor), ptr(state_spect_tmp1.tensor)
        # rk4_step0
        call("b2_step_euler", h, S, dt / 6, ptr(tendencies_0.tensor), ptr(diss), acc, nvar, stream_ptr())
        call("b2_step_euler", h, S, dt / 2, ptr(tendencies_0.tensor), ptr(diss2), tmp1, nvar, stream_ptr())
        tendencies_1 = sim.tendencies_nonlin(state_spect_tmp1, old=tendencies_0)
        call("b2_rk4_step1", h, S, acc, tmp1, ptr(tendencies_1.tensor), ptr(diss2), dt, nvar, stream_ptr())
        tendencies_2 = sim.tendencies_nonlin(state_spect_tmp1, old=tendencies_1)
        call("b2_rk4_step2", h, S, acc, tmp1, ptr(tendencies_2.tensor), ptr(diss), ptr(diss2), dt, nvar,
             stream_ptr())
        tendencies_3 = sim.tendencies_nonlin(state_spect_tmp1, old=tendencies_2)
        call("b2_rk4_step3", h, S, acc, ptr(tendencies_3.tensor), dt, nvar, stream_ptr())


class TimeSteppingPseudoSpectralStratB200(TimeSteppingPseudoSpectralB200):
    """``solvers/ns2d/strat/time_stepping.py:19-186``: the ns2d.strat time stepper adds the time-step
    limits of the internal gravity waves to the advective CFL rule."""

    def _init_compute_time_step(self):
        super()._init_compute_time_step()
        from math import pi

        oper = self.sim.oper
        N = float(self.params.N)
        self.coef_deltat_dispersion_relation = 1.0
        self.coef_group = getattr(self.params.time_stepping, "cfl_coef_group", 1.0)
        self.coef_phase = 1.0
        KX, KZ = oper.KX, oper.KY
        K_not0 = torch.sqrt(oper.K2_not0)
        # compute_dispersion_relation (ns2d/strat/solver.py:215-225)
        freq_disp_relation = float((N * (KX / K_not0)).max().item())
        self.deltat_dispersion_relation = self.coef_deltat_dispersion_relation * (2.0 * pi / freq_disp_relation)
        if self.coef_group:  # _compute_time_increment_group_and_phase (:109-139)
            cg_kx = (N / K_not0) * (KZ**2 / K_not0**2)
            cg_kz = (-N / K_not0) * ((KX / K_not0) * (KZ / K_not0))
            freq_group = float(cg_kx.max().item()) / oper.deltax + float(cg_kz.max().item()) / oper.deltay
            freq_phase = float((N * (KX / K_not0**2)).max().item()) / oper.deltax
            self.deltat_group_vel = self.coef_group / freq_group
            self.deltat_phase_vel = self.coef_phase / freq_phase
        if self.params.forcing.enable:  # _compute_time_increment_forcing (:100-107)
            self.deltat_f = 1.0 / (self.params.forcing.forcing_rate ** (1.0 / 3))

    def compute_time_increment_CLF(self):
        """_compute_time_increment_CFL_uxuyb (:141-186)."""
        get_var = self.sim.state.get_var
        oper = self.sim.oper
        freq_CFL = self._max_abs(get_var("ux")) / oper.deltax + self._max_abs(get_var("uy")) / oper.deltay
        deltat_CFL = self.CFL / freq_CFL if freq_CFL > 0 else self.deltat_max
        if not self.coef_group:
            maybe_new_dt = min(deltat_CFL, self.deltat_dispersion_relation, self.deltat_max)
        else:
            maybe_new_dt = min(deltat_CFL, self.deltat_dispersion_relation, self.deltat_group_vel, self.deltat_max)
        if self.params.forcing.enable:
            maybe_new_dt = min(maybe_new_dt, self.deltat_f)
        normalize_diff = abs(self.deltat - maybe_new_dt) / maybe_new_dt
        if normalize_diff > 0.02:
            self.deltat = maybe_new_dt
