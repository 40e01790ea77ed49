"""Forcing on the GPU path (mirror of ``/root/reference/fluidsim/base/forcing/base.py:160-200`` and of
the forcing makers of ``base/forcing/specific.py``).

The reference adds ``self.forcing.get_forcing()`` (a ``state_spect``-shaped ``forcing_fft``) to the
tendencies of every RK stage (``solvers/ns3d/solver.py:243-244``, ``ns2d/solver.py:190-191``); the
forcing itself is recomputed once per time step (``base/time_stepping/base.py:212-213``).  fluidsim's
makers force a small set of low-wavenumber modes (``SpecificForcingPseudoSpectralCoarse``: a coarse
grid of ``(2 nkmax_forcing)^3`` modes at most), so the fused path receives ``forcing_fft`` as a SPARSE
list (C ABI ``b2_set_forcing_sparse``) and adds it in front of the RK epilogue: no field-sized pass.

Makers implemented here:

* ``in_script``        user function returning ``{key: K array}`` (``InScriptForcingPseudoSpectral``,
                       specific.py:98-134; physical-space variant through ``oper.fft``)
* ``proportional``     ``f = -(forcing_rate / sum |v_c|^2 ...)``: see ``Proportional`` (specific.py:381-425)
* ``tcrandom``         time-correlated random forcing normalised to a constant injection rate
                       (``TimeCorrelatedRandomPseudoSpectral`` + ``NormalizedForcing`` with
                       ``type_normalize = "2nd_degree_eq"``, specific.py:428-870), host RNG, forced
                       shell ``nkmin_forcing <= |k| / deltak <= nkmax_forcing``

The random / normalised makers work on the forced shell only (a few hundred modes gathered from the
GPU state each step), which is what the reference does on its coarse operator.
"""

import math
import types

import numpy as np
import torch

from ._lib import call, ptr
from .setofvariables import SetOfVariables

MAX_FORCED_MODES = 1 << 22


class SpecificForcing:
    tag = "specific"

    def __init__(self, sim):
        self.sim = sim
        self.oper = sim.oper
        self.params = sim.params
        self.forcing_fft = SetOfVariables(like=sim.state.state_spect, info="forcing_fft", value=0.0)

    # what the keys of a forcing dict mean for this solver: state keys only on the GPU path
    def _set_from_dict(self, kwargs):
        keys = self.sim.state.keys_state_spect
        self.forcing_fft.initialize(0.0)
        for key, value in kwargs.items():
            if key not in keys:
                raise ValueError(f"forcing key {key!r}: the GPU path forces state variables {keys}")
            if isinstance(value, np.ndarray):
                value = torch.from_numpy(value)
            self.forcing_fft.tensor[keys.index(key)].copy_(value.to(self.oper.device))


class InScriptForcingPseudoSpectral(SpecificForcing):
    """specific.py:98-134."""

    tag = "in_script"

    def __init__(self, sim):
        super().__init__(sim)
        self.is_initialized = False

    def compute(self):
        obj = self.compute_forcing_fft_each_time()
        if isinstance(obj, dict):
            kwargs = obj
        else:
            if self.params.forcing.key_forced is None:
                raise ValueError("params.forcing.key_forced must be initialized.")
            kwargs = {self.params.forcing.key_forced: obj}
        self._set_from_dict(kwargs)

    def compute_forcing_fft_each_time(self):
        obj = self.compute_forcing_each_time()
        fft = self.oper.fft
        if isinstance(obj, dict):
            return {key: fft(value) for key, value in obj.items()}
        if self.params.forcing.key_forced is None:
            raise ValueError("params.forcing.key_forced must be initialized.")
        return {self.params.forcing.key_forced: fft(obj)}

    def compute_forcing_each_time(self):
        return self.oper.create_arrayX(value=0)

    def monkeypatch_compute_forcing_fft_each_time(self, func):
        self.compute_forcing_fft_each_time = types.MethodType(func, self)
        self.is_initialized = True

    def monkeypatch_compute_forcing_each_time(self, func):
        self.compute_forcing_each_time = types.MethodType(func, self)
        self.is_initialized = True


class ShellForcing(SpecificForcing):
    """Common part of the makers acting on the shell kmin_forcing <= |k| <= kmax_forcing
    (SpecificForcingPseudoSpectralCoarse, specific.py:137-345, restricted to the forced modes)."""

    def __init__(self, sim):
        super().__init__(sim)
        pf = self.params.forcing
        if pf.nkmax_forcing < pf.nkmin_forcing:
            raise ValueError(
                f"params.forcing.nkmax_forcing = {pf.nkmax_forcing} < "
                f"params.forcing.nkmin_forcing = {pf.nkmin_forcing}"
            )
        oper = self.oper
        self.kmax_forcing = oper.deltak * pf.nkmax_forcing
        self.kmin_forcing = oper.deltak * pf.nkmin_forcing
        self.forcing_rate = float(pf.forcing_rate)
        K = torch.sqrt(oper.K2)
        cond = (K <= self.kmax_forcing) & (K >= self.kmin_forcing) & (oper.where_dealiased == 0)
        self.ind_forcing = torch.nonzero(cond.reshape(-1)).reshape(-1)  # linear K-field indices
        self.nb_forced_modes = int(self.ind_forcing.numel())
        if not self.nb_forced_modes:
            raise ValueError("0 modes forced.")
        # r2c-aware weights of sum_wavenumbers on the forced modes
        nk = oper.shapeK_loc[-1]
        ikx = (self.ind_forcing % nk).cpu().numpy()
        w = np.full(self.nb_forced_modes, 2.0)
        w[ikx == 0] = 1.0
        if oper.nx % 2 == 0:
            w[ikx == nk - 1] = 1.0
        self._w = w
        self._K = {name: getattr(oper, name).reshape(-1)[self.ind_forcing].cpu().numpy()
                   for name in ("Kx", "Ky", "Kz") if hasattr(oper, name)}
        keys = sim.state.keys_state_spect
        self.nforced_vars = 1 if sim.ndim == 2 else 3

    def _gather_state(self):
        """state_spect on the forced modes -> (nvar, nmodes) complex numpy array."""
        S = self.sim.state.state_spect.tensor
        flat = S.reshape(S.shape[0], -1)
        return flat[: self.nforced_vars, self.ind_forcing].cpu().numpy()

    def _scatter(self, f_modes):
        self.forcing_fft.initialize(0.0)
        flat = self.forcing_fft.tensor.reshape(self.forcing_fft.tensor.shape[0], -1)
        flat[: self.nforced_vars, self.ind_forcing] = torch.from_numpy(np.ascontiguousarray(f_modes)).to(
            self.oper.device)

    def _sumk(self, a):
        return float((self._w * a).sum())

    def _project(self, f):
        """Leray projection of a (3, nmodes) forcing: keeps the forced velocity solenoidal."""
        if self.sim.ndim == 2:
            return f
        Kx, Ky, Kz = self._K["Kx"], self._K["Ky"], self._K["Kz"]
        K2 = Kx * Kx + Ky * Ky + Kz * Kz
        K2 = np.where(K2 == 0, 1e-14, K2)
        tmp = (Kx * f[0] + Ky * f[1] + Kz * f[2]) / K2
        return np.stack([f[0] - Kx * tmp, f[1] - Ky * tmp, f[2] - Kz * tmp])


class Proportional(ShellForcing):
    """Forcing proportional to the forced variable (specific.py:381-425):
    f = alpha v on the forced shell, alpha = (sqrt(1 + dt P / Z) - 1) / dt, Z = sum' |v|^2 / 2."""

    tag = "proportional"

    def compute(self):
        v = self._gather_state()
        Z = self._sumk((np.abs(v) ** 2).sum(0)) / 2.0
        deltat = float(self.sim.time_stepping.deltat)
        alpha = (math.sqrt(1 + deltat * self.forcing_rate / Z) - 1.0) / deltat if Z > 0 else 0.0
        self._scatter(alpha * v)


class TimeCorrelatedRandomPseudoSpectral(ShellForcing):
    """Time-correlated random forcing with constant energy injection rate
    (TimeCorrelatedRandomPseudoSpectral, specific.py:768-870, on top of NormalizedForcing with
    type_normalize = "2nd_degree_eq", specific.py:587-726).

    Two random fields f0, f1 are renewed every ``time_correlation``; in between the raw forcing moves
    from f0 to f1 with a raised-cosine weight (forcingc_from_f0f1).  The raw forcing f_r is rescaled,
    f = R f_r, R a root of  a R^2 + b R + c = 0,  a = dt/2 sum'|f_r|^2,  b = sum' Re(conj(v) f_r),
    c = -forcing_rate (normalize_forcingc_2nd_degree_eq), which fixes the energy injected over one
    time step.  Restated on the forced shell for the velocity VECTOR (the reference normalises each
    forced key of its coarse state separately and draws from fluidfft's create_arrayK_random [EXT],
    so the random streams differ; the injection-rate identity is what the tests pin)."""

    tag = "tcrandom"

    def __init__(self, sim):
        super().__init__(sim)
        pf = self.params.forcing
        tc = getattr(pf, "tcrandom", None)
        time_correlation = getattr(tc, "time_correlation", "based_on_forcing_rate") if tc is not None else \
            "based_on_forcing_rate"
        if time_correlation == "based_on_forcing_rate":
            self.period_change_f0f1 = self.forcing_rate ** (-1.0 / 3)
        else:
            self.period_change_f0f1 = float(time_correlation)
        seed = getattr(pf, "random_seed", None)
        self.rng = np.random.default_rng(0 if seed is None else seed)
        self.t_last_change = float(sim.time_stepping.t) if hasattr(sim, "time_stepping") else 0.0
        self.forcing0 = self._raw()
        self.forcing1 = self._raw()

    def _raw(self):
        shape = (self.nforced_vars, self.nb_forced_modes)
        f = self.rng.uniform(-1, 1, shape) + 1j * self.rng.uniform(-1, 1, shape)
        return self._project(f)

    def _from_f0f1(self, t):
        """forcingc_raw_each_time + forcingc_from_f0f1 (specific.py:826-870)."""
        if t - self.t_last_change >= self.period_change_f0f1:
            self.t_last_change = t
            self.forcing0 = self.forcing1
            self.forcing1 = self._raw()
        omega = math.pi / self.period_change_f0f1
        deltaf = self.forcing1 - self.forcing0
        return self.forcing1 - 0.5 * (math.cos((t - self.t_last_change) * omega) + 1) * deltaf

    def coef_normalization_from_abc(self, a, b, c):
        """specific.py:679-726."""
        try:
            alpha1, alpha2 = np.roots([a, b, c])
        except ValueError:
            return 0.0
        norm = getattr(self.params.forcing, "normalized", None)
        which_root = getattr(norm, "which_root", "minabs") if norm is not None else "minabs"
        if which_root == "minabs":
            return alpha2 if abs(alpha2) < abs(alpha1) else alpha1
        if which_root == "first":
            return alpha1
        if which_root == "second":
            return alpha2
        if which_root == "positive":
            return alpha2 if alpha2 > 0.0 else alpha1
        raise ValueError("Not sure how to choose which root to normalize forcing with.")

    def compute(self):
        """NormalizedForcing.compute + normalize_forcingc_2nd_degree_eq (specific.py:472-509, 587-677)."""
        ts = self.sim.time_stepping
        f_r = self._from_f0f1(float(ts.t))
        v = self._gather_state()
        deltat = float(ts.deltat)
        a = deltat / 2 * self._sumk((np.abs(f_r) ** 2).sum(0))
        b = self._sumk((v.conj() * f_r).real.sum(0))
        c = -self.forcing_rate
        self._scatter(float(np.real(self.coef_normalization_from_abc(a, b, c))) * f_r)


FORCING_CLASSES = {
    cls.tag: cls for cls in (InScriptForcingPseudoSpectral, Proportional, TimeCorrelatedRandomPseudoSpectral)
}


class ForcingB200:
    """base/forcing/base.py:88-200 (ForcingBasePseudoSpectral) for the GPU Simul classes."""

    def __init__(self, sim):
        params = sim.params
        self.type_forcing = params.forcing.type
        if self.type_forcing not in FORCING_CLASSES:
            raise ValueError("Wrong value for params.forcing.type: " + str(self.type_forcing))
        self.sim = sim
        self.forcing_maker = FORCING_CLASSES[self.type_forcing](sim)
        self._t_last_computed = -math.inf
        self._idx = self._val = None

    def __call__(self, key):
        return self.get_forcing().get_var(key)

    def compute(self):
        time = self.sim.time_stepping.t
        if time > self._t_last_computed:
            self.forcing_maker.compute()
            self._t_last_computed = time
            self._push_sparse()

    def get_forcing(self):
        return self.forcing_maker.forcing_fft

    def is_initialized(self):
        return getattr(self.forcing_maker, "is_initialized", True)

    def _push_sparse(self):
        """forcing_fft -> sparse (index, value) list held by the plan (fused path)."""
        sim = self.sim
        F = self.forcing_maker.forcing_fft.tensor
        nv = 1 if sim.ndim == 2 else 3
        if F.shape[0] > nv and bool((F[nv:] != 0).any()):
            raise NotImplementedError("forcing of the buoyancy is not implemented on the fused GPU path")
        flat = F[:nv].reshape(nv, -1)
        nz = (flat != 0).any(dim=0)
        idx = torch.nonzero(nz).reshape(-1)
        if idx.numel() > MAX_FORCED_MODES:
            raise NotImplementedError(
                f"{idx.numel()} forced modes: the fused path takes sparse low-wavenumber forcing only"
            )
        mask = sim.oper.where_dealiased
        if mask is not None and idx.numel() and bool((mask.reshape(-1)[idx] != 0).any()):
            raise ValueError("the forcing acts on dealiased modes")
        self._idx = idx.to(torch.int64).contiguous()
        self._val = flat[:, idx].contiguous()
        call("b2_set_forcing_sparse", sim.oper.plan.handle, int(idx.numel()), ptr(self._idx),
             ptr(torch.view_as_real(self._val)), nv)
