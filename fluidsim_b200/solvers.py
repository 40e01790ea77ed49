"""GPU ``Simul`` classes for ns3d, ns3d.strat, ns3d.bouss, ns2d, ns2d.strat and ns2d.bouss.

Host-side mirror of ``/root/reference/fluidsim/solvers/ns3d/solver.py:57-263``,
``solvers/ns3d/strat/solver.py:57-216`` and ``solvers/ns2d/solver.py:73-194`` reduced to the hot
path: construction order Operators -> State -> TimeStepping (``base/solvers/base.py:117-223``),
``tendencies_nonlin(state_spect=None, old=None)``, ``project_state_spect``,
``compute_freq_diss``.  Same names, same argument meaning, same exceptions; arrays are CUDA
tensors and all arithmetic is done by libb200spectral kernels.
"""

import numpy as np
import torch

from ._lib import SOLVER_IDS, call, lib, ptr, stream_ptr
from .operators import OperatorsPseudoSpectral2D, OperatorsPseudoSpectral3D, vector_product
from .params import create_default_params
from .setofvariables import SetOfVariables
from .state import StateNS2D, StateNS2DStrat, StateNS3D, StateNS3DStrat
from .time_stepping import TimeSteppingPseudoSpectralB200, TimeSteppingPseudoSpectralStratB200


class PhysFieldsB200:
    """``sim.output.phys_fields.save()`` (base/output/phys_fields.py:130-202)."""

    def __init__(self, output):
        self.output = output

    def save(self, state_phys=None, params=None, particular_attr=None):
        from .checkpoint import save_state_phys

        out = self.output
        return save_state_phys(out.sim, out.path_run, out.name_run, particular_attr=particular_attr)


class OutputB200:
    """The slice of ``OutputBase`` the checkpoint needs: ``path_run``, ``name_run``, ``phys_fields``."""

    def __init__(self, sim):
        self.sim = sim
        self.name_run = f"{sim.short_name}_b200"
        self.path_run = getattr(getattr(sim.params, "output", None), "path_run", None) or "."
        self.phys_fields = PhysFieldsB200(self)


class SimulBasePseudoSpectralB200:
    short_name = None
    ndim = 3
    # False: the solver runs on the operator-level kernels only (no fused b2_time_step / b2_tendencies)
    supports_fused = True
    Operators = OperatorsPseudoSpectral3D
    State = StateNS3D
    TimeStepping = TimeSteppingPseudoSpectralB200

    @classmethod
    def create_default_params(cls):
        return create_default_params(cls.short_name)

    def __init__(self, params, fused=None):
        self.params = params
        self.is_forcing_enabled = bool(getattr(params.forcing, "enable", False))
        self.oper = self.Operators(params)
        self.state = self.State(self)
        self._init_projection()
        self._fused_buffers = None
        self._fused_mask = None
        # dealias-pruned transforms: used for a step only when the state is known to be dealiased
        # (true after every step; reset by state.mark_spect_modified()).
        self.use_pruning = True
        self._state_dealiased = False
        self.time_stepping = self.TimeStepping(self, fused=fused)
        # base/solvers/base.py:190-195: the forcing object comes after the time stepper
        self.forcing = None
        if self.is_forcing_enabled:
            from .forcing import ForcingB200

            self.forcing = ForcingB200(self)
        # base/solvers/base.py:196-216: output object, then the initial fields.  Only the checkpoint
        # pieces of both exist here: output.phys_fields.save() and
        # init_fields.type = "from_file"; other initial conditions are set through the state container.
        self.output = OutputB200(self)
        init = getattr(params, "init_fields", None)
        if init is not None and getattr(init, "type", None) == "from_file":
            from .checkpoint import load_state_phys

            load_state_phys(self, init.from_file.path)

    def _init_projection(self):
        pass

    # ---- fused path plumbing -------------------------------------------------------------------------
    def _physics_args(self):
        p = self.params
        f = getattr(p, "f", None)
        return (
            float(p.nu_2), float(p.nu_4), float(p.nu_8), float(p.nu_m4),
            0 if f is None else 1, 0.0 if f is None else float(f),
            0.0 if self.short_name == "ns3d.bouss" else float(getattr(p, "N", 0.0)), float(getattr(p, "beta", 0.0)),
        )

    def _mask_for_fused(self):
        oper = self.oper
        if self.ndim == 2 and not oper._has_to_dealiase:
            return None
        return oper.where_dealiased

    def _ensure_fused_buffers(self):
        """Allocate acc / stage / work buffers (caller-owned by the C ABI) and push the physics."""
        mask = self._mask_for_fused()
        if self._fused_buffers is not None and self._fused_mask is mask:
            return
        oper = self.oper
        h = oper.plan.handle
        import ctypes as C

        nwork, nvar = C.c_int(), C.c_int()
        call("b2_work_fields", h, SOLVER_IDS[self.short_name], C.byref(nwork), C.byref(nvar))
        # ns3d: the raw transform outputs are aliased with the stage buffer (b2_set_aliasing), so the
        # work buffer holds the 3 vorticity fields only: 12 K fields per GPU instead of 15
        import os

        alias = self.short_name == "ns3d" and os.environ.get("B2_NOALIAS", "0") in ("0", "")
        if self._fused_buffers is None:
            mk = lambda n: torch.empty((n,) + tuple(oper.shapeK_loc), dtype=torch.complex128, device=oper.device)
            self._fused_buffers = (mk(nvar.value), mk(nvar.value), mk(3 if alias else nwork.value))
        acc, stage, work = self._fused_buffers
        self._fused_mask = mask
        call("b2_set_physics", h, SOLVER_IDS[self.short_name], *self._physics_args(), ptr(mask))
        call("b2_set_buffers", h, ptr(acc), ptr(stage), ptr(work))
        call("b2_set_aliasing", h, 1 if alias else 0)
        if self.ndim == 3:
            call("b2_set_no_vz_kz0", h, 1 if getattr(self, "no_vz_kz0", False) else 0)
            call("b2_set_projection", h, int(getattr(self, "_projection_id", 0)))

    def mask_modified(self):
        """Call after editing ``oper.where_dealiased`` in place: the kept ranges of the pruned
        transforms are recomputed and the state is re-checked before the next fused step."""
        self._fused_mask = None
        self._state_dealiased = False

    def tendencies_nonlin_fused(self, state_spect=None, old=None):
        """N(state_spect) through the fused kernels (C ABI ``b2_tendencies``)."""
        self._ensure_fused_buffers()
        call("b2_set_pruning", self.oper.plan.handle, 0)  # arbitrary input: no assumption on zeros
        src = self.state.state_spect if state_spect is None else state_spect
        tendencies_fft = SetOfVariables(like=self.state.state_spect, info="tendencies_nonlin") if old is None else old
        call("b2_tendencies", self.oper.plan.handle, ptr(src.tensor), ptr(tendencies_fft.tensor), stream_ptr())
        return tendencies_fft

    # ---- linear term (base/solvers/pseudo_spect.py:134-191) -----------------------------------------
    def compute_freq_diss(self):
        p = self.params
        oper = self.oper
        K2 = oper.K2
        f_d = p.nu_2 * K2 if p.nu_2 > 0 else torch.zeros_like(K2)
        if p.nu_4 > 0.0:
            f_d = f_d + p.nu_4 * K2**2
        if p.nu_8 > 0.0:
            f_d = f_d + p.nu_8 * K2**4
        if p.nu_m4 != 0.0:
            f_d_hypo = p.nu_m4 / oper.K2_not0**2
            if self.ndim == 2:
                f_d_hypo[0, 0] = f_d_hypo[0, 1]
            else:
                f_d_hypo[0, 0, 0] = f_d_hypo[0, 0, 1]
        else:
            f_d_hypo = 0.0
        return f_d, f_d_hypo


class SimulNS3D(SimulBasePseudoSpectralB200):
    """solvers/ns3d/solver.py:57-263."""

    short_name = "ns3d"

    def __init__(self, params, fused=None):
        super().__init__(params, fused=fused)
        oper = self.oper
        # solvers/ns3d/state.py:46-52 (allocated lazily: only the unfused path needs them)
        self._fields_tmp = None
        self._fields_spect_tmp = None

    def _init_projection(self):
        """solver.py:147-174: only the default projection is on the GPU path."""
        self.no_vz_kz0 = bool(getattr(self.params, "no_vz_kz0", False))
        if self.no_vz_kz0:  # solver.py:153-157
            self.where_kz_0 = self.oper.Kz.abs() == 0.0
        projection = getattr(self.params, "projection", None)
        if projection is None:
            self._projector = self.oper.project_perpk3d
            self._projection_id = 0
        elif projection in ("toroidal", "vortical"):
            self._projector = self.oper.project_toroidal
            self._projection_id = 1
        elif projection == "poloidal":
            self._projector = self.oper.project_poloidal
            self._projection_id = 2
        else:
            raise ValueError(f"No known projection for params.projection = {projection}")

    @property
    def fields_tmp(self):
        if self._fields_tmp is None:
            self._fields_tmp = tuple(self.oper.create_arrayX() for _ in range(6))
        return self._fields_tmp

    @property
    def fields_spect_tmp(self):
        if self._fields_spect_tmp is None:
            self._fields_spect_tmp = tuple(self.oper.create_arrayK() for _ in range(3))
        return self._fields_spect_tmp

    def _modif_omegafft_with_f(self, omegax_fft, omegay_fft, omegaz_fft):
        omegaz_fft[0, 0, 0] += self.params.f

    def tendencies_nonlin(self, state_spect=None, old=None):
        """solver.py:180-253, executed with operator-level kernels (any grid size)."""
        oper = self.oper
        ifft_as_arg = oper.ifft_as_arg
        ifft_as_arg_destroy = oper.ifft_as_arg_destroy
        fft_as_arg = oper.fft_as_arg
        spect_get_var = (self.state.state_spect if state_spect is None else state_spect).get_var
        vx_fft = spect_get_var("vx_fft")
        vy_fft = spect_get_var("vy_fft")
        vz_fft = spect_get_var("vz_fft")
        omegax_fft, omegay_fft, omegaz_fft = self.fields_spect_tmp
        oper.rotfft_from_vecfft_outin(vx_fft, vy_fft, vz_fft, omegax_fft, omegay_fft, omegaz_fft)
        if self.params.f is not None:
            self._modif_omegafft_with_f(omegax_fft, omegay_fft, omegaz_fft)
        omegax, omegay, omegaz = self.fields_tmp[3:6]
        ifft_as_arg_destroy(omegax_fft, omegax)
        ifft_as_arg_destroy(omegay_fft, omegay)
        ifft_as_arg_destroy(omegaz_fft, omegaz)
        if state_spect is None:
            vx = self.state.state_phys.get_var("vx")
            vy = self.state.state_phys.get_var("vy")
            vz = self.state.state_phys.get_var("vz")
        else:
            vx, vy, vz = self.fields_tmp[0:3]
            ifft_as_arg(vx_fft, vx)
            ifft_as_arg(vy_fft, vy)
            ifft_as_arg(vz_fft, vz)
        fx, fy, fz = vector_product(vx, vy, vz, omegax, omegay, omegaz)
        if old is None:
            tendencies_fft = SetOfVariables(like=self.state.state_spect, info="tendencies_nonlin")
        else:
            tendencies_fft = old
        fft_as_arg(fx, tendencies_fft.get_var("vx_fft"))
        fft_as_arg(fy, tendencies_fft.get_var("vy_fft"))
        fft_as_arg(fz, tendencies_fft.get_var("vz_fft"))
        self._extra_tendencies(tendencies_fft, spect_get_var, state_spect, vx, vy, vz)
        if self.is_forcing_enabled:  # solver.py:243-244
            tendencies_fft += self.forcing.get_forcing()
        self.project_state_spect(tendencies_fft)
        self.oper.dealiasing(tendencies_fft)
        return tendencies_fft

    def _extra_tendencies(self, tendencies_fft, spect_get_var, state_spect, vx, vy, vz):
        pass

    def project_state_spect(self, state_spect):
        """solver.py:255-263."""
        self._projector(
            state_spect.get_var("vx_fft"), state_spect.get_var("vy_fft"), state_spect.get_var("vz_fft")
        )
        if self.no_vz_kz0:
            state_spect.get_var("vz_fft")[self.where_kz_0] = 0.0
            if "b_fft" in state_spect.keys:
                state_spect.get_var("b_fft")[self.where_kz_0] = 0.0


class SimulNS3DStrat(SimulNS3D):
    """solvers/ns3d/strat/solver.py:57-216."""

    short_name = "ns3d.strat"
    State = StateNS3DStrat
    _N_coupling = None  # None: params.N (ns3d.bouss overrides with 0)

    def _extra_tendencies(self, tendencies_fft, spect_get_var, state_spect, vx, vy, vz):
        """strat/solver.py:198-211: fz += b ; fb = -div(v b) - N^2 vz."""
        oper = self.oper
        b_fft = spect_get_var("b_fft")
        vz_fft = spect_get_var("vz_fft")
        fz_fft = tendencies_fft.get_var("vz_fft")
        call("b2_add_inplace", ptr(fz_fft), ptr(b_fft), fz_fft.numel(), stream_ptr())
        if state_spect is None:
            b = self.state.state_phys.get_var("b")
        else:
            b = self.fields_tmp[3]
            oper.ifft_as_arg(b_fft, b)
        div_vb_fft = oper.div_vb_fft_from_vb(vx, vy, vz, b)
        N = self._N_coupling if self._N_coupling is not None else float(self.params.N)
        call("b2_compute_fb_fft", ptr(div_vb_fft), float(N), ptr(vz_fft), div_vb_fft.numel(), stream_ptr())
        tendencies_fft.set_var("b_fft", div_vb_fft)


class SimulNS3DBouss(SimulNS3DStrat):
    """solvers/ns3d/bouss/solver.py:99-175: the stratified solver without the background
    stratification term, fb = -div(v b) (fz += b kept).  Same kernels with N = 0."""

    short_name = "ns3d.bouss"
    _N_coupling = 0.0


class SimulNS2D(SimulBasePseudoSpectralB200):
    """solvers/ns2d/solver.py:73-194."""

    short_name = "ns2d"
    ndim = 2
    Operators = OperatorsPseudoSpectral2D
    State = StateNS2D

    def __init__(self, params, fused=None):
        super().__init__(params, fused=fused)
        self._fields_tmp = None

    @property
    def fields_tmp(self):
        if self._fields_tmp is None:
            self._fields_tmp = tuple(self.oper.create_arrayX() for _ in range(4))
        return self._fields_tmp

    def tendencies_nonlin(self, state_spect=None, old=None):
        oper = self.oper
        ifft_as_arg_destroy = oper.oper_fft.ifft_as_arg_destroy
        if state_spect is None:
            rot_fft = self.state.state_spect.get_var("rot_fft")
            ux = self.state.state_phys.get_var("ux")
            uy = self.state.state_phys.get_var("uy")
        else:
            rot_fft = state_spect.get_var("rot_fft")
            ux_fft, uy_fft = oper.vecfft_from_rotfft(rot_fft)
            ux, uy = self.fields_tmp[0:2]
            ifft_as_arg_destroy(ux_fft, ux)
            ifft_as_arg_destroy(uy_fft, uy)
        px_rot_fft, py_rot_fft = oper.gradfft_from_fft(rot_fft)
        px_rot, py_rot = self.fields_tmp[2:4]
        ifft_as_arg_destroy(px_rot_fft, px_rot)
        ifft_as_arg_destroy(py_rot_fft, py_rot)
        # compute_Frot (solver.py:34-38), result in px_rot's buffer
        call("b2_compute_frot", ptr(ux), ptr(uy), ptr(px_rot), ptr(py_rot), float(self.params.beta),
             ptr(px_rot), px_rot.numel(), stream_ptr())
        if old is None:
            tendencies_fft = SetOfVariables(like=self.state.state_spect)
        else:
            tendencies_fft = old
        Frot_fft = tendencies_fft.get_var("rot_fft")
        oper.fft_as_arg(px_rot, Frot_fft)
        oper.dealiasing(Frot_fft)
        if self.params.forcing.enable:  # ns2d/solver.py:190-191
            tendencies_fft += self.forcing.get_forcing()
        return tendencies_fft


class SimulNS2DStrat(SimulNS2D):
    """solvers/ns2d/strat/solver.py:59-181 -- state (rot_fft, b_fft); runs on the operator-level kernels
    (FFT passes, gradfft / vecfft, one elementwise product kernel)."""

    short_name = "ns2d.strat"
    State = StateNS2DStrat
    TimeStepping = TimeSteppingPseudoSpectralStratB200  # ns2d/strat/solver.py:53-56
    supports_fused = False
    _bouss = 0

    @property
    def fields_tmp(self):
        if self._fields_tmp is None:  # field_tmp0..5 of ns2d/state.py:43-46 + strat/state.py:50-54
            self._fields_tmp = tuple(self.oper.create_arrayX() for _ in range(6))
        return self._fields_tmp

    def tendencies_nonlin(self, state_spect=None, old=None):
        """strat/solver.py:71-181 (bouss/solver.py:65-173 with ``_bouss``)."""
        oper = self.oper
        ifft_as_arg = oper.ifft_as_arg
        ifft_as_arg_destroy = oper.oper_fft.ifft_as_arg_destroy
        fft_as_arg = oper.fft_as_arg
        if old is None:
            tendencies_fft = SetOfVariables(like=self.state.state_spect)
        else:
            tendencies_fft = old
        f_rot_fft = tendencies_fft.get_var("rot_fft")
        f_b_fft = tendencies_fft.get_var("b_fft")
        if state_spect is None:
            rot_fft = self.state.state_spect.tensor[0]
            b_fft = self.state.state_spect.tensor[1]
            ux = self.state.state_phys.get_var("ux")
            uy = self.state.state_phys.get_var("uy")
        else:
            rot_fft = state_spect.get_var("rot_fft")
            b_fft = state_spect.get_var("b_fft")
            ux_fft, uy_fft = oper.vecfft_from_rotfft(rot_fft)
            ux, uy = self.fields_tmp[0:2]
            ifft_as_arg_destroy(ux_fft, ux)
            ifft_as_arg_destroy(uy_fft, uy)
        px_rot_fft, py_rot_fft = oper.gradfft_from_fft(rot_fft)
        px_b_fft, py_b_fft = oper.gradfft_from_fft(b_fft)
        px_rot, py_rot, px_b, py_b = self.fields_tmp[2:6]
        ifft_as_arg_destroy(px_rot_fft, px_rot)
        ifft_as_arg_destroy(py_rot_fft, py_rot)
        ifft_as_arg(px_b_fft, px_b)  # px_b_fft is used again below
        ifft_as_arg_destroy(py_b_fft, py_b)
        # f_rot in px_rot's buffer, f_b in py_rot's (px_b is an input of both expressions)
        call("b2_tendencies_ns2d_buoyancy", ptr(ux), ptr(uy), ptr(px_rot), ptr(py_rot), ptr(px_b), ptr(py_b),
             float(getattr(self.params, "N", 0.0)), int(self._bouss), ptr(px_rot), ptr(py_rot), px_rot.numel(),
             stream_ptr())
        fft_as_arg(py_rot, f_b_fft)
        fft_as_arg(px_rot, f_rot_fft)
        if not self._bouss:  # strat/solver.py:156
            call("b2_add_inplace", ptr(f_rot_fft), ptr(px_b_fft), f_rot_fft.numel(), stream_ptr())
        oper.dealiasing(tendencies_fft)
        if self.params.forcing.enable:
            tendencies_fft += self.forcing.get_forcing()
        return tendencies_fft


class SimulNS2DBouss(SimulNS2DStrat):
    """solvers/ns2d/bouss/solver.py:60-173."""

    short_name = "ns2d.bouss"
    TimeStepping = TimeSteppingPseudoSpectralB200  # bouss keeps the base stepper (bouss/solver.py:30-57)
    _bouss = 1


SIMUL_CLASSES = {
    "ns3d": SimulNS3D, "ns3d.strat": SimulNS3DStrat, "ns3d.bouss": SimulNS3DBouss,
    "ns2d": SimulNS2D, "ns2d.strat": SimulNS2DStrat, "ns2d.bouss": SimulNS2DBouss,
}


def make_simul(solver, params, fused=None):
    return SIMUL_CLASSES[solver](params, fused=fused)
