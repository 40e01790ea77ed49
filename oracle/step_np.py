"""Self-contained numpy restatement of the reference hot path (TEST INFRASTRUCTURE).

Travels to the GPU box (no /root/reference there).  Each function cites the reference
lines it follows; paths are relative to /root/reference/fluidsim.  Pinned here, in the
build container, against the reference's own modules executed through
``oracle.refshim`` (``tests/test_oracle.py``, bit for bit, every solver and every time scheme;
fixtures in ``tests/golden``
made by ``tests/golden/make_golden.py``).  The fluidfft layer underneath
(``oracle.fluidfft_np``) is a restatement of an absent third-party dependency:
parity of that layer is UNPINNED (see ``oracle/__init__.py``).
"""

from math import pi

import numpy as np

from .fluidfft_np import (
    OperatorsPseudoSpectral2D,
    OperatorsPseudoSpectral3D,
    SetOfVariables,
    vector_product,
)


# ------------------------------------------------------------------ operators (fluidsim level)
def reinit_truncation(oper, truncation_shape, ndim):
    """operators/base.py:47-72 (``OperatorBase._reinit_truncation``)."""
    if truncation_shape == "cubic":
        return
    kmax = oper.coef_dealiasing * oper.deltakx * oper.nx / 2
    if truncation_shape == "spherical":
        oper.where_dealiased = np.array(oper.K2 >= kmax**2, dtype=np.uint8)
    elif truncation_shape == "no_multiple_aliases":
        # operators3d.py:280-288 / operators2d.py:195-198
        ax = abs(oper.Kx if ndim == 3 else oper.KX) >= 2 / 3 * oper.deltakx * oper.nx / 2
        ay = abs(oper.Ky if ndim == 3 else oper.KY) >= 2 / 3 * oper.deltaky * oper.ny / 2
        if ndim == 3:
            az = abs(oper.Kz) >= 2 / 3 * oper.deltakz * oper.nz / 2
            where = (ax & ay) | (ay & az) | (az & ax)
        else:
            where = ax & ay
        if oper.coef_dealiasing:
            where |= oper.K2 >= kmax**2
        oper.where_dealiased = np.array(where, dtype=np.uint8)
    else:
        raise ValueError(
            'truncation_shape must be "cubic", "spherical" or "no_multiple_aliases"'
        )


def dealiasing_setofvar(sov, where_dealiased):
    """operators3d.py:38-59,74-76 (numpy variant ``dealiasing_setofvar_numpy``)."""
    nz = np.nonzero(where_dealiased)
    for i in range(sov.shape[0]):
        sov[i][nz] = 0.0


# ------------------------------------------------------------------ time-stepping kernels
def step_Euler(state_spect, dt, tendencies, diss, output):
    """base/time_stepping/pseudo_spect.py:51-56."""
    output[:] = (state_spect + dt * tendencies) * diss
    return output


def step_like_RK2(state_spect, dt, tendencies, diss, diss2):
    """base/time_stepping/pseudo_spect.py:64-68."""
    state_spect[:] = state_spect * diss + dt * diss2 * tendencies


class OracleSim:
    """ns3d | ns3d.strat | ns3d.bouss | ns2d | ns2d.strat | ns2d.bouss simulation object reduced to
    its hot path.

    Mirrors the construction order of ``base/solvers/base.py:117-223`` and the per-step
    sequence of ``solvers/ns3d/time_stepping.py:8-20`` /
    ``base/time_stepping/pseudo_spect.py:236-243``.
    """

    def __init__(
        self,
        solver,
        nx,
        ny,
        nz=None,
        Lx=2 * pi,
        Ly=2 * pi,
        Lz=2 * pi,
        nu_2=0.0,
        nu_4=0.0,
        nu_8=0.0,
        nu_m4=0.0,
        coef_dealiasing=2.0 / 3,
        truncation_shape="cubic",
        type_time_scheme="RK4",
        deltat0=1e-2,
        N=1.0,
        f=None,
        beta=0.0,
        no_vz_kz0=False,
        projection=None,
        nb_pairs=1,
        nb_steps_compute_new_pair=None,
    ):
        self.solver = solver
        # params.time_stepping.phaseshift_random (pseudo_spect.py:159-167)
        self.nb_pairs = nb_pairs
        self.nb_steps_compute_new_pair = nb_steps_compute_new_pair
        self._phaseshift = None
        self.nu_2, self.nu_4, self.nu_8, self.nu_m4 = nu_2, nu_4, nu_8, nu_m4
        self.N, self.f, self.beta = N, f, beta
        self.no_vz_kz0 = bool(no_vz_kz0)
        self.projection = projection
        self.deltat = float(deltat0)
        self.scheme = type_time_scheme
        self.it = 0
        self.t = 0.0
        # forcing_fft of the current step (same shape as state_spect) or None; the reference adds
        # `self.forcing.get_forcing()` to the tendencies (solvers/ns3d/solver.py:243-244)
        self.forcing_fft = None
        if solver in ("ns2d", "ns2d.strat", "ns2d.bouss"):
            self.ndim = 2
            self.oper = OperatorsPseudoSpectral2D(nx, ny, Lx, Ly, coef_dealiasing=coef_dealiasing)
            self.oper.Lx, self.oper.Ly = self.oper.lx, self.oper.ly
            if solver == "ns2d":
                keys_spect = ["rot_fft"]
                keys_phys = ["ux", "uy", "rot"]
            else:  # ns2d/strat/state.py:27-47, ns2d/bouss/state.py:27-38
                keys_spect = ["rot_fft", "b_fft"]
                keys_phys = ["ux", "uy", "rot", "b"]
        elif solver in ("ns3d", "ns3d.strat", "ns3d.bouss"):
            self.ndim = 3
            self.oper = OperatorsPseudoSpectral3D(
                nx, ny, nz, Lx, Ly, Lz, coef_dealiasing=coef_dealiasing
            )
            keys_phys = ["vx", "vy", "vz"] + (["b"] if solver in ("ns3d.strat", "ns3d.bouss") else [])
            keys_spect = [k + "_fft" for k in keys_phys]
        else:
            raise ValueError(solver)
        reinit_truncation(self.oper, truncation_shape, self.ndim)
        oper = self.oper
        # base/state.py:243-248, 63-68
        self.state_spect = SetOfVariables(
            keys=keys_spect, shape_variable=oper.shapeK_loc, dtype=np.complex128, info="state_spect"
        )
        self.state_phys = SetOfVariables(
            keys=keys_phys, shape_variable=oper.shapeX_loc, dtype=np.float64, info="state_phys"
        )
        self.state_spect[:] = 0
        self.state_phys[:] = 0
        # ns2d/state.py:43-46 (+ field_tmp4/5 of ns2d/strat/state.py:50-54), ns3d/state.py:46-52
        n_tmp = 6 if self.ndim == 3 or solver != "ns2d" else 4
        self.fields_tmp = tuple(np.empty(oper.shapeX_loc) for _ in range(n_tmp))
        self.fields_spect_tmp = tuple(
            np.empty(oper.shapeK_loc, dtype=np.complex128) for _ in range(3)
        )
        # pseudo_spect.py:179-189, 103-152
        self.freq_lin = self.compute_freq_diss()
        self.exact = np.exp(-self.deltat * self.freq_lin)
        self.exact2 = np.exp(-self.deltat / 2 * self.freq_lin)
        self._state_spect_tmp = np.empty_like(self.state_spect)
        self._state_spect_tmp1 = np.empty_like(self.state_spect)
        if type_time_scheme.endswith("_random"):  # pseudo_spect.py:210-216
            self._init_phaseshift_random()

    # ------------------------------------------------------------------ linear term
    def compute_freq_diss(self):
        """base/solvers/pseudo_spect.py:134-191."""
        oper = self.oper
        if self.nu_2 > 0:
            f_d = self.nu_2 * oper.K2
        else:
            f_d = np.zeros_like(oper.K2)
        if self.nu_4 > 0.0:
            f_d += self.nu_4 * oper.K2**2
        if self.nu_8 > 0.0:
            f_d += self.nu_8 * oper.K8
        if self.nu_m4 != 0.0:
            K2_not0 = np.copy(oper.K2)
            K2_not0[(0,) * self.ndim] = 1e-14
            f_d_hypo = self.nu_m4 / K2_not0**2
            if self.ndim == 2:
                f_d_hypo[0, 0] = f_d_hypo[0, 1]
            else:
                f_d_hypo[0, 0, 0] = f_d_hypo[0, 0, 1]
        else:
            f_d_hypo = 0.0
        return f_d + f_d_hypo

    def set_deltat(self, dt):
        """ExactLinearCoefs.compute, pseudo_spect.py:122-141."""
        self.deltat = float(dt)
        self.exact = np.exp(-dt * self.freq_lin)
        self.exact2 = np.exp(-dt / 2 * self.freq_lin)

    # ------------------------------------------------------------------ state sync
    def statephys_from_statespect(self):
        """base/state.py:326-332; ns2d/state.py:95-106."""
        oper = self.oper
        if self.ndim == 2:
            rot_fft = self.state_spect.get_var("rot_fft")
            ux_fft, uy_fft = oper.vecfft_from_rotfft(rot_fft)
            oper.ifft_as_arg(rot_fft, self.state_phys.get_var("rot"))
            oper.ifft_as_arg(ux_fft, self.state_phys.get_var("ux"))
            oper.ifft_as_arg(uy_fft, self.state_phys.get_var("uy"))
            if self.solver != "ns2d":  # ns2d/strat/state.py:137-151, ns2d/bouss/state.py:96-110
                oper.ifft_as_arg(self.state_spect.get_var("b_fft"), self.state_phys.get_var("b"))
        else:
            for ik in range(self.state_spect.nvar):
                oper.ifft_as_arg(self.state_spect[ik].view(np.ndarray), self.state_phys[ik].view(np.ndarray))

    def set_state_spect(self, arr):
        self.state_spect[...] = arr
        self.statephys_from_statespect()

    # ------------------------------------------------------------------ nonlinear terms
    def project_state_spect(self, state_spect):
        """solvers/ns3d/solver.py:255-263 with params.projection (:158-174)."""
        vx_fft, vy_fft, vz_fft = (state_spect.get_var(k) for k in ("vx_fft", "vy_fft", "vz_fft"))
        if self.projection is None:
            self.oper.project_perpk3d(vx_fft, vy_fft, vz_fft)
        elif self.projection in ("toroidal", "vortical"):
            self._project_toroidal(vx_fft, vy_fft, vz_fft)
        elif self.projection == "poloidal":
            self._project_poloidal(vx_fft, vy_fft, vz_fft)
        else:
            raise ValueError(f"No known projection for params.projection = {self.projection}")
        if self.no_vz_kz0:  # solver.py:260-263
            where_kz_0 = np.abs(self.oper.Kz) == 0.0
            state_spect.get_var("vz_fft")[where_kz_0] = 0.0
            if "b_fft" in state_spect.keys:
                state_spect.get_var("b_fft")[where_kz_0] = 0.0

    def _project_toroidal(self, vx_fft, vy_fft, vz_fft):
        """operators/operators3d.py:911-958."""
        oper = self.oper
        Kh_square_nozero = oper.Kx**2 + oper.Ky**2
        Kh_square_nozero[Kh_square_nozero == 0] = 1e-14
        tmp = np.sqrt(1.0 / Kh_square_nozero)
        cos_phi_k = oper.Kx * tmp
        sin_phi_k = oper.Ky * tmp
        tmp = -sin_phi_k * vx_fft + cos_phi_k * vy_fft
        vx_fft[...] = -sin_phi_k * tmp
        vy_fft[...] = cos_phi_k * tmp
        vz_fft[...] = 0.0

    def _project_poloidal(self, vx_fft, vy_fft, vz_fft):
        """operators/operators3d.py:788-856."""
        oper = self.oper
        Kh_square = oper.Kx**2 + oper.Ky**2
        K_square_nozero = Kh_square + oper.Kz**2
        Kh_square_nozero = Kh_square.copy()
        Kh_square_nozero[Kh_square_nozero == 0] = 1e-14
        K_square_nozero[K_square_nozero == 0] = 1e-14
        inv_Kh_square_nozero = 1.0 / Kh_square_nozero
        inv_K_square_nozero = 1.0 / K_square_nozero
        cos_theta_k = oper.Kz * np.sqrt(inv_K_square_nozero)
        sin_theta_k = np.sqrt(Kh_square * inv_K_square_nozero)
        cos_phi_k = oper.Kx * np.sqrt(inv_Kh_square_nozero)
        sin_phi_k = oper.Ky * np.sqrt(inv_Kh_square_nozero)
        tmp = cos_theta_k * cos_phi_k * vx_fft + cos_theta_k * sin_phi_k * vy_fft - sin_theta_k * vz_fft
        vx_fft[...] = cos_theta_k * cos_phi_k * tmp
        vy_fft[...] = cos_theta_k * sin_phi_k * tmp
        vz_fft[...] = -sin_theta_k * tmp

    def dealiasing(self, thing):
        """operators3d.py:336-342; operators2d.py:200-220."""
        oper = self.oper
        if self.ndim == 2 and not oper._has_to_dealiase:
            return
        if isinstance(thing, SetOfVariables):
            dealiasing_setofvar(thing, oper.where_dealiased)
        else:
            thing[np.nonzero(oper.where_dealiased)] = 0.0

    def tendencies_nonlin(self, state_spect=None, old=None):
        if self.solver == "ns2d":
            return self._tendencies_ns2d(state_spect, old)
        if self.ndim == 2:
            return self._tendencies_ns2d_buoyancy(state_spect, old)
        return self._tendencies_ns3d(state_spect, old)

    def _tendencies_ns3d(self, state_spect=None, old=None):
        """solvers/ns3d/solver.py:180-253; strat extras solvers/ns3d/strat/solver.py:138-216."""
        oper = self.oper
        strat = self.solver in ("ns3d.strat", "ns3d.bouss")
        get = (self.state_spect if state_spect is None else state_spect).get_var
        vx_fft, vy_fft, vz_fft = get("vx_fft"), get("vy_fft"), get("vz_fft")
        omegax_fft, omegay_fft, omegaz_fft = self.fields_spect_tmp
        oper.rotfft_from_vecfft_outin(vx_fft, vy_fft, vz_fft, omegax_fft, omegay_fft, omegaz_fft)
        if self.f is not None:
            omegaz_fft[0, 0, 0] += self.f  # solver.py:176-178
        omegax, omegay, omegaz = self.fields_tmp[3:6]
        oper.ifft_as_arg_destroy(omegax_fft, omegax)
        oper.ifft_as_arg_destroy(omegay_fft, omegay)
        oper.ifft_as_arg_destroy(omegaz_fft, omegaz)
        if state_spect is None:
            vx = self.state_phys.get_var("vx")
            vy = self.state_phys.get_var("vy")
            vz = self.state_phys.get_var("vz")
        else:
            vx, vy, vz = self.fields_tmp[0:3]
            oper.ifft_as_arg(vx_fft, vx)
            oper.ifft_as_arg(vy_fft, vy)
            oper.ifft_as_arg(vz_fft, vz)
        fx, fy, fz = vector_product(vx, vy, vz, omegax, omegay, omegaz)
        if old is None:
            tendencies_fft = SetOfVariables(like=self.state_spect, info="tendencies_nonlin")
        else:
            tendencies_fft = old
        oper.fft_as_arg(fx, tendencies_fft.get_var("vx_fft"))
        oper.fft_as_arg(fy, tendencies_fft.get_var("vy_fft"))
        oper.fft_as_arg(fz, tendencies_fft.get_var("vz_fft"))
        if strat:
            b_fft = get("b_fft")
            fz_fft = tendencies_fft.get_var("vz_fft")
            fz_fft += b_fft  # strat/solver.py:198
            if state_spect is None:
                b = self.state_phys.get_var("b")
            else:
                b = self.fields_tmp[3]
                oper.ifft_as_arg(b_fft, b)
            div_vb_fft = oper.div_vb_fft_from_vb(vx, vy, vz, b)
            if self.solver == "ns3d.bouss":  # bouss/solver.py:166
                fb_fft = -div_vb_fft
            else:  # compute_fb_fft, strat/solver.py:29-33
                fb_fft = -div_vb_fft - self.N**2 * vz_fft
            tendencies_fft.set_var("b_fft", fb_fft)
        if self.forcing_fft is not None:  # solvers/ns3d/solver.py:243-244
            tendencies_fft += self.forcing_fft
        self.project_state_spect(tendencies_fft)
        self.dealiasing(tendencies_fft)
        return tendencies_fft

    def _tendencies_ns2d(self, state_spect=None, old=None):
        """solvers/ns2d/solver.py:111-194; compute_Frot :34-38."""
        oper = self.oper
        if state_spect is None:
            rot_fft = self.state_spect.get_var("rot_fft")
            ux = self.state_phys.get_var("ux")
            uy = self.state_phys.get_var("uy")
        else:
            rot_fft = state_spect.get_var("rot_fft")
            ux_fft, uy_fft = oper.vecfft_from_rotfft(rot_fft)
            ux, uy = self.fields_tmp[0:2]
            oper.ifft_as_arg(ux_fft, ux)
            oper.ifft_as_arg(uy_fft, uy)
        px_rot_fft, py_rot_fft = oper.gradfft_from_fft(rot_fft)
        px_rot, py_rot = self.fields_tmp[2:4]
        oper.ifft_as_arg(px_rot_fft, px_rot)
        oper.ifft_as_arg(py_rot_fft, py_rot)
        if self.beta == 0:
            Frot = -ux * px_rot - uy * py_rot
        else:
            Frot = -ux * px_rot - uy * (py_rot + self.beta)
        if old is None:
            tendencies_fft = SetOfVariables(like=self.state_spect)
        else:
            tendencies_fft = old
        Frot_fft = tendencies_fft.get_var("rot_fft")
        oper.fft_as_arg(Frot, Frot_fft)
        self.dealiasing(Frot_fft)
        if self.forcing_fft is not None:  # solvers/ns2d/solver.py:190-191 (after the dealiasing)
            tendencies_fft += self.forcing_fft
        return tendencies_fft

    def _tendencies_ns2d_buoyancy(self, state_spect=None, old=None):
        """solvers/ns2d/strat/solver.py:71-181 (tendencies_nonlin_ns2dstrat :21-27) and
        solvers/ns2d/bouss/solver.py:65-173 (tendencies_nonlin_ns2dbouss :21-27)."""
        oper = self.oper
        if old is None:
            tendencies_fft = SetOfVariables(like=self.state_spect)
        else:
            tendencies_fft = old
        f_rot_fft = tendencies_fft.get_var("rot_fft")
        f_b_fft = tendencies_fft.get_var("b_fft")
        if state_spect is None:
            rot_fft = self.state_spect.get_var("rot_fft")
            b_fft = self.state_spect.get_var("b_fft")
            ux = self.state_phys.get_var("ux")
            uy = self.state_phys.get_var("uy")
        else:
            rot_fft = state_spect.get_var("rot_fft")
            b_fft = state_spect.get_var("b_fft")
            ux_fft, uy_fft = oper.vecfft_from_rotfft(rot_fft)
            ux, uy = self.fields_tmp[0:2]
            oper.ifft_as_arg(ux_fft, ux)
            oper.ifft_as_arg(uy_fft, uy)
        px_rot_fft, py_rot_fft = oper.gradfft_from_fft(rot_fft)
        px_b_fft, py_b_fft = oper.gradfft_from_fft(b_fft)
        px_rot, py_rot, px_b, py_b = self.fields_tmp[2:6]
        oper.ifft_as_arg(px_rot_fft, px_rot)
        oper.ifft_as_arg(py_rot_fft, py_rot)
        oper.ifft_as_arg(px_b_fft, px_b)
        oper.ifft_as_arg(py_b_fft, py_b)
        if self.solver == "ns2d.strat":
            f_rot = -ux * px_rot - uy * py_rot
            f_b = -ux * px_b - uy * py_b - self.N**2 * uy
        else:
            f_rot = -ux * px_rot - uy * py_rot + px_b
            f_b = -ux * px_b - uy * py_b
        oper.fft_as_arg(f_b, f_b_fft)
        oper.fft_as_arg(f_rot, f_rot_fft)
        if self.solver == "ns2d.strat":
            f_rot_fft += px_b_fft  # strat/solver.py:156
        self.dealiasing(tendencies_fft)
        if self.forcing_fft is not None:
            tendencies_fft += self.forcing_fft
        return tendencies_fft

    # ------------------------------------------------------------------ schemes
    def _time_step_RK2(self):
        """base/time_stepping/pseudo_spect.py:469-517."""
        dt = self.deltat
        diss, diss2 = self.exact, self.exact2
        state_spect = self.state_spect
        tendencies_0 = self.tendencies_nonlin()
        state_spect_12 = self._state_spect_tmp
        step_Euler(state_spect, dt / 2, tendencies_0, diss2, output=state_spect_12)
        tendencies_12 = self.tendencies_nonlin(state_spect_12, old=tendencies_0)
        step_like_RK2(state_spect, dt, tendencies_12, diss, diss2)

    def _time_step_RK4(self):
        """base/time_stepping/pseudo_spect.py:798-984."""
        dt = self.deltat
        diss, diss2 = self.exact, self.exact2
        state_spect = self.state_spect
        tendencies_0 = self.tendencies_nonlin()
        state_spect_tmp1 = self._state_spect_tmp1
        state_spect_tmp = step_Euler(state_spect, dt / 6, tendencies_0, diss, output=self._state_spect_tmp)
        state_spect_12_approx1 = step_Euler(
            state_spect, dt / 2, tendencies_0, diss2, output=state_spect_tmp1
        )
        tendencies_1 = self.tendencies_nonlin(state_spect_12_approx1, old=tendencies_0)
        state_spect_12_approx2 = state_spect_tmp1
        state_spect_tmp[:] += dt / 3 * diss2 * tendencies_1  # :938
        state_spect_12_approx2[:] = state_spect * diss2 + dt / 2 * tendencies_1  # :939-941
        tendencies_2 = self.tendencies_nonlin(state_spect_12_approx2, old=tendencies_1)
        state_spect_1_approx = state_spect_tmp1
        state_spect_tmp[:] += dt / 3 * diss2 * tendencies_2  # :968
        state_spect_1_approx[:] = state_spect * diss + dt * diss2 * tendencies_2  # :969-971
        tendencies_3 = self.tendencies_nonlin(state_spect_1_approx, old=tendencies_2)
        state_spect[:] = state_spect_tmp + dt / 6 * tendencies_3  # :984

    # ---- Euler / trapezoid / phase-shifting schemes (pseudo_spect.py:245-468, 519-796)
    def _get_phaseshift(self):
        """pseudo_spect.py:281-300."""
        if self._phaseshift is None:
            oper = self.oper
            if self.ndim == 2:
                phase = 0.5 * (oper.deltax * oper.KX + oper.deltay * oper.KY)
            else:
                phase = 0.5 * (oper.deltax * oper.Kx + oper.deltay * oper.Ky + oper.deltaz * oper.Kz)
            self._phaseshift = np.exp(1j * phase)
        return self._phaseshift

    def _get_phases_random(self):
        """operators3d.py:1128-1146 / operators2d.py:618-631 (Python's `random`, not numpy's)."""
        from random import uniform

        oper = self.oper
        if self.ndim == 3:
            alpha_x, alpha_y, alpha_z = tuple(uniform(-0.5, 0.5) for _ in range(3))
            beta_x = alpha_x + 0.5 if alpha_x < 0 else alpha_x - 0.5
            beta_y = alpha_y + 0.5 if alpha_y < 0 else alpha_y - 0.5
            beta_z = alpha_z + 0.5 if alpha_z < 0 else alpha_z - 0.5
            phase_alpha = (
                alpha_x * oper.deltax * oper.Kx + alpha_y * oper.deltay * oper.Ky + alpha_z * oper.deltaz * oper.Kz
            )
            phase_beta = beta_x * oper.deltax * oper.Kx + beta_y * oper.deltay * oper.Ky + beta_z * oper.deltaz * oper.Kz
        else:
            alpha_x, alpha_y = tuple(uniform(-0.5, 0.5) for _ in range(2))
            beta_x = alpha_x + 0.5 if alpha_x < 0 else alpha_x - 0.5
            beta_y = alpha_y + 0.5 if alpha_y < 0 else alpha_y - 0.5
            phase_alpha = alpha_x * oper.deltax * oper.KX + alpha_y * oper.deltay * oper.KY
            phase_beta = beta_x * oper.deltax * oper.KX + beta_y * oper.deltay * oper.KY
        return phase_alpha, phase_beta

    def _init_phaseshift_random(self):
        """pseudo_spect.py:302-328."""
        if self.nb_steps_compute_new_pair is None:
            self.nb_steps_compute_new_pair = 2 if self.nb_pairs == 1 else 4 * self.nb_pairs
        self._index_phaseshift = 1
        self._previous_index_pair = 0
        self._previous_index_flip = 0
        self._pairs_phaseshift = []
        for _ in range(self.nb_pairs):
            phase_alpha, phase_beta = self._get_phases_random()
            self._pairs_phaseshift.append((np.exp(1j * phase_alpha), np.exp(1j * phase_beta)))

    def _get_phaseshift_random(self):
        """pseudo_spect.py:330-372.  On the step that renews the oldest pair the reference re-binds
        (alpha, beta) to that pair's arrays and overwrites them in place (compute_phaseshift_terms,
        :91-100): the step then uses the NEW pair, in (alpha, beta) order."""
        from random import randint

        nb_pairs, nb_steps = self.nb_pairs, self.nb_steps_compute_new_pair
        if nb_pairs == 1 and nb_steps == 1:
            phaseshift_alpha, phaseshift_beta = self._pairs_phaseshift[0]
        elif nb_pairs == 1 and nb_steps == 2:
            pair = self._pairs_phaseshift[0]
            if self._index_phaseshift == 1:
                phaseshift_alpha, phaseshift_beta = pair
            else:
                phaseshift_beta, phaseshift_alpha = pair
        else:
            index_pair = randint(0, nb_pairs - 1)
            pair = self._pairs_phaseshift[index_pair]
            index_flip = randint(0, 1)
            if index_pair == self._previous_index_pair and index_flip == self._previous_index_flip:
                index_flip = 0 if index_flip else 1
            self._previous_index_pair = index_pair
            self._previous_index_flip = index_flip
            if index_flip:
                phaseshift_alpha, phaseshift_beta = pair
            else:
                phaseshift_beta, phaseshift_alpha = pair
        if self._index_phaseshift == nb_steps:
            self._index_phaseshift = 1
            phase_alpha, phase_beta = self._get_phases_random()
            phaseshift_alpha, phaseshift_beta = self._pairs_phaseshift.pop(0)
            phaseshift_alpha[:] = np.exp(1j * phase_alpha)
            phaseshift_beta[:] = np.exp(1j * phase_beta)
            self._pairs_phaseshift.append((phaseshift_alpha, phaseshift_beta))
        else:
            self._index_phaseshift += 1
        return phaseshift_alpha, phaseshift_beta

    def _like_state(self, arr):
        out = SetOfVariables(like=self.state_spect)
        out[...] = arr
        return out

    def _time_step_Euler(self):
        """pseudo_spect.py:245-279."""
        tendencies_0 = self.tendencies_nonlin()
        self.state_spect[:] = (self.state_spect + self.deltat * tendencies_0) * self.exact

    def _time_step_Euler_phaseshift(self):
        """pseudo_spect.py:374-420."""
        state_spect = self.state_spect
        tendencies_0 = self.tendencies_nonlin()
        phaseshift = self._get_phaseshift()
        tendencies_shifted = self.tendencies_nonlin(self._like_state(phaseshift * state_spect)) / phaseshift
        tendencies_dealiased = 0.5 * (tendencies_0 + tendencies_shifted)
        state_spect[:] = (state_spect + self.deltat * tendencies_dealiased) * self.exact

    def _time_step_Euler_phaseshift_random(self):
        """pseudo_spect.py:422-467."""
        state_spect = self.state_spect
        phaseshift_alpha, phaseshift_beta = self._get_phaseshift_random()
        tendencies_alpha = self.tendencies_nonlin(self._like_state(phaseshift_alpha * state_spect)) / phaseshift_alpha
        tendencies_beta = self.tendencies_nonlin(self._like_state(phaseshift_beta * state_spect)) / phaseshift_beta
        tendencies_dealiased = 0.5 * (tendencies_alpha + tendencies_beta)
        state_spect[:] = (state_spect + self.deltat * tendencies_dealiased) * self.exact

    def _time_step_RK2_trapezoid(self):
        """pseudo_spect.py:519-569."""
        dt = self.deltat
        diss = self.exact
        state_spect = self.state_spect
        tendencies_0 = self.tendencies_nonlin()
        state_spect_1 = step_Euler(state_spect, dt, tendencies_0, diss, output=self._state_spect_tmp)
        tendencies_1 = self.tendencies_nonlin(state_spect_1)
        state_spect[:] = (state_spect + dt / 2 * tendencies_0) * diss + dt / 2 * tendencies_1

    def _time_step_RK2_phaseshift(self):
        """pseudo_spect.py:571-638."""
        dt = self.deltat
        diss, diss2 = self.exact, self.exact2
        state_spect = self.state_spect
        tendencies_0 = self.tendencies_nonlin()
        state_spect_1 = step_Euler(state_spect, dt, tendencies_0, diss, output=self._state_spect_tmp)
        phaseshift = self._get_phaseshift()
        tendencies_1_shift = self.tendencies_nonlin(self._like_state(phaseshift * state_spect_1))
        tendencies_d = self._state_spect_tmp
        tendencies_d[:] = 0.5 * (tendencies_0 + tendencies_1_shift / phaseshift)
        step_like_RK2(state_spect, dt, tendencies_d, diss, diss2)

    def _time_step_RK2_phaseshift_random(self):
        """pseudo_spect.py:640-709.  Both tendencies calls write their result over their input
        (``old=state_spect_shift``), which changes what ns3d.strat computes for fb_fft."""
        dt = self.deltat
        diss, diss2 = self.exact, self.exact2
        phaseshift_alpha, phaseshift_beta = self._get_phaseshift_random()
        state_spect = self.state_spect
        state_spect_shift = self._like_state(phaseshift_alpha * state_spect)
        tendencies_0_shift = self.tendencies_nonlin(state_spect_shift, old=state_spect_shift)
        tendencies_0_shift /= phaseshift_alpha  # div_inplace
        tendencies_0 = tendencies_0_shift
        state_spect_1 = step_Euler(state_spect, dt, tendencies_0, diss, output=self._state_spect_tmp)
        state_spect_1_shift = self._like_state(phaseshift_beta * state_spect_1)
        tendencies_1_shift = self.tendencies_nonlin(state_spect_1_shift, old=state_spect_1_shift)
        tendencies_d = self._state_spect_tmp
        tendencies_d[:] = 0.5 * (tendencies_0 + tendencies_1_shift / phaseshift_beta)  # mean_with_phaseshift
        step_like_RK2(state_spect, dt, tendencies_d, diss, diss2)

    def _time_step_RK2_phaseshift_exact(self):
        """pseudo_spect.py:711-796 (the two shifted evaluations alias output and input, see above)."""
        dt = self.deltat
        diss, diss2 = self.exact, self.exact2
        phaseshift = self._get_phaseshift()
        state_spect = self.state_spect
        tmp0 = SetOfVariables(like=state_spect)
        tendencies_0 = self.tendencies_nonlin(state_spect, old=tmp0)
        state_spect_shift = self._like_state(phaseshift * state_spect)
        tendencies_0_shift = self.tendencies_nonlin(state_spect_shift, old=state_spect_shift)
        tendencies_d0 = 0.5 * (tendencies_0 + tendencies_0_shift / phaseshift)
        state_spect_1 = step_Euler(state_spect, dt, tendencies_d0, diss, output=self._state_spect_tmp)
        tendencies_1 = self.tendencies_nonlin(state_spect_1, old=tmp0)
        state_spect_shift = self._like_state(phaseshift * state_spect_1)
        tendencies_1_shift = self.tendencies_nonlin(state_spect_shift, old=state_spect_shift)
        tendencies_d = 0.5 * (tendencies_d0 + 0.5 * (tendencies_1 + tendencies_1_shift / phaseshift))
        step_like_RK2(state_spect, dt, tendencies_d, diss, diss2)

    def one_time_step(self):
        """ns3d/time_stepping.py:8-20 (3-D) / pseudo_spect.py:236-243 (2-D) + base.py:243-244."""
        schemes = ("RK4", "RK2", "Euler", "Euler_phaseshift", "Euler_phaseshift_random", "RK2_trapezoid",
                   "RK2_phaseshift", "RK2_phaseshift_random", "RK2_phaseshift_exact")
        if self.scheme not in schemes:  # pseudo_spect.py:191-224
            raise ValueError(f'Problem name time_scheme ("{self.scheme}")')
        getattr(self, "_time_step_" + self.scheme)()
        if self.ndim == 3:
            self.project_state_spect(self.state_spect)
        self.dealiasing(self.state_spect)
        self.statephys_from_statespect()
        if np.isnan(np.sum(self.state_spect[0])):
            raise ValueError(f"nan at it = {self.it}, t = {self.t:.4f}")
        self.t += self.deltat
        self.it += 1
        return self.state_spect

    # ------------------------------------------------------------------ CFL (base.py:320-354)
    def compute_time_increment_CFL(self, cfl=1.0, deltat_max=0.2):
        """base/time_stepping/base.py:320-354 (RK4: CFL = 1.0, :259-260)."""
        oper = self.oper
        if self.ndim == 3:
            vx, vy, vz = (self.state_phys.get_var(k) for k in ("vx", "vy", "vz"))
            freq = (
                np.abs(vx).max() / oper.deltax
                + np.abs(vy).max() / oper.deltay
                + np.abs(vz).max() / oper.deltaz
            )
        else:
            ux, uy = self.state_phys.get_var("ux"), self.state_phys.get_var("uy")
            freq = np.abs(ux).max() / oper.deltax + np.abs(uy).max() / oper.deltay
        deltat_CFL = cfl / freq if freq > 0 else deltat_max
        deltat_wanted = min(deltat_CFL, deltat_max)
        if self.solver == "ns2d.strat":  # _compute_time_increment_CFL_uxuyb, ns2d/strat/time_stepping.py:149-186
            lim = self.strat_time_increments()
            if not self.cfl_coef_group:
                deltat_wanted = min(deltat_CFL, lim["dispersion_relation"], deltat_max)
            else:
                deltat_wanted = min(deltat_CFL, lim["dispersion_relation"], lim["group_vel"], deltat_max)
            if self.forcing_rate is not None:
                deltat_wanted = min(deltat_wanted, 1.0 / (self.forcing_rate ** (1.0 / 3)))
        if abs(self.deltat - deltat_wanted) / deltat_wanted > 0.02:  # base.py:350-354
            self.set_deltat(deltat_wanted)
        return self.deltat

    # ns2d.strat: params.time_stepping.cfl_coef_group (None by default) and params.forcing.forcing_rate
    # when the forcing is enabled (ns2d/strat/time_stepping.py:24-29,100-107)
    cfl_coef_group = None
    forcing_rate = None

    def strat_time_increments(self):
        """Wave time-step limits of ns2d.strat (ns2d/strat/time_stepping.py:57-98,109-147;
        compute_dispersion_relation: ns2d/strat/solver.py:215-225)."""
        oper = self.oper
        N = self.N
        KX, KZ, K_not0 = oper.KX, oper.KY, oper.K_not0
        out = {"dispersion_relation": 1.0 * (2.0 * pi / (N * (KX / K_not0)).max())}
        if self.cfl_coef_group:
            cg_kx = (N / K_not0) * (KZ**2 / K_not0**2)
            cg_kz = (-N / K_not0) * ((KX / K_not0) * (KZ / K_not0))
            freq_group = cg_kx.max() / oper.deltax + cg_kz.max() / oper.deltay
            cp = N * (KX / K_not0**2)
            freq_phase = cp.max() / oper.deltax
            out["group_vel"] = self.cfl_coef_group / freq_group
            out["phase_vel"] = 1.0 / freq_phase
        return out

    # ------------------------------------------------------------------ observables
    def compute_energy(self):
        """ns3d/output/__init__.py:101-117 ; ns2d energy = sum' |rot|^2/K2 / 2."""
        oper = self.oper
        if self.ndim == 3:
            e = 0.0
            for k in ("vx_fft", "vy_fft", "vz_fft"):
                e += oper.sum_wavenumbers(0.5 * np.abs(self.state_spect.get_var(k)) ** 2)
            return e
        rot_fft = self.state_spect.get_var("rot_fft")
        return oper.sum_wavenumbers(0.5 * np.abs(rot_fft) ** 2 / oper.K2_not0)

    def compute_enstrophy(self):
        oper = self.oper
        if self.ndim == 2:
            return oper.sum_wavenumbers(0.5 * np.abs(self.state_spect.get_var("rot_fft")) ** 2)
        ox, oy, oz = oper.rotfft_from_vecfft(*(self.state_spect.get_var(k) for k in ("vx_fft", "vy_fft", "vz_fft")))
        return oper.sum_wavenumbers(0.5 * (np.abs(ox) ** 2 + np.abs(oy) ** 2 + np.abs(oz) ** 2))

    def compute_spectrum3d(self):
        e = sum(
            0.5 * np.abs(self.state_spect.get_var(k)) ** 2 for k in ("vx_fft", "vy_fft", "vz_fft")
        )
        return self.oper.compute_3dspectrum(e)

    # ------------------------------------------------------------------ initial fields
    def compute_spatial_means(self):
        """The state reductions of SpatialMeansNS3D._save_one_time
        (solvers/ns3d/output/spatial_means.py:23-43): E, Ex, Ey, Ez, epsK, epsK_hypo, epsK4, epsK8."""
        oper = self.oper
        s = np.array(self.state_spect)
        nrj = [0.5 * np.abs(s[i]) ** 2 for i in range(3)]
        energy_fft = nrj[0] + nrj[1] + nrj[2]
        sw = oper.sum_wavenumbers
        K2 = oper.K2
        f_d = self.nu_2 * K2 if self.nu_2 > 0 else np.zeros_like(K2)
        if self.nu_4 > 0:
            f_d = f_d + self.nu_4 * K2**2
        if self.nu_8 > 0:
            f_d = f_d + self.nu_8 * K2**4
        out = dict(Ex=sw(nrj[0]), Ey=sw(nrj[1]), Ez=sw(nrj[2]), epsK=sw(f_d * 2 * energy_fft))
        out["E"] = out["Ex"] + out["Ey"] + out["Ez"]
        if self.nu_m4 != 0.0:
            K2n = K2.copy()
            K2n[0, 0, 0] = K2[0, 0, 1]
            out["epsK_hypo"] = sw(self.nu_m4 / K2n**2 * 2 * energy_fft)
        else:
            out["epsK_hypo"] = 0.0
        out["epsK4"] = sw(self.nu_4 * K2**2 * 2 * energy_fft) if self.nu_4 > 0 else 0.0
        out["epsK8"] = sw(self.nu_8 * K2**4 * 2 * energy_fft) if self.nu_8 > 0 else 0.0
        return out

    def compute_spectra(self):
        """SpectraNS3D.compute (solvers/ns3d/output/spectra.py:15-60): per-component 1-D and 3-D spectra."""
        oper = self.oper
        s = np.array(self.state_spect)
        out = {}
        for i, key in enumerate(("vx", "vy", "vz")):
            nrj = 0.5 * np.abs(s[i]) ** 2
            out[key + "_kx"], out[key + "_ky"], out[key + "_kz"] = oper.compute_1dspectra(nrj)
            out[key] = oper.compute_3dspectrum(nrj)
        out["E"] = out["vx"] + out["vy"] + out["vz"]
        for ax in ("kx", "ky", "kz"):
            out["E_" + ax] = out["vx_" + ax] + out["vy_" + ax] + out["vz_" + ax]
        return out

    def init_noise(self, velo_max=1.0, length=None, seed=42):
        """ns3d/init_fields.py:110-196 ; ns2d/init_fields.py:57-108."""
        oper = self.oper

        def H_smooth(x, delta):
            return (1.0 + np.tanh(2 * pi * x / delta)) / 2.0

        if self.ndim == 3:
            lambda0 = oper.Lx / 4.0 if length is None else length
            np.random.seed(seed)
            vv = [np.random.random(oper.shapeX_loc) - 0.5 for _ in range(3)]
            vv_fft = []
            for vi in vv:
                vi_fft = oper.fft(vi)
                vi_fft[0, 0, 0] = 0.0
                vv_fft.append(vi_fft)
            oper.project_perpk3d(*vv_fft)
            for vi_fft in vv_fft:
                self.dealiasing(vi_fft)
            k0 = 2 * pi / lambda0
            delta_k0 = 1.0 * k0
            K = np.sqrt(oper.K2)
            vv_fft = [vi_fft * H_smooth(k0 - K, delta_k0) for vi_fft in vv_fft]
            vv = [oper.ifft(ui_fft) for ui_fft in vv_fft]
            vmax = np.sqrt(vv[0] ** 2 + vv[1] ** 2 + vv[2] ** 2).max()
            vv = [velo_max * vi / vmax for vi in vv]
            fields = [oper.fft(vi) for vi in vv]
            if self.solver in ("ns3d.strat", "ns3d.bouss"):
                lambda0 = min(oper.Lx, oper.Ly, oper.Lz) / 4.0 if length is None else length
                k0 = 2 * pi / lambda0
                field = np.random.random(oper.shapeX_loc)
                field_fft = oper.fft(field)
                field_fft[0, 0, 0] = 0.0
                field_fft *= H_smooth(k0 - K, 1.0 * k0)
                oper.ifft_as_arg(field_fft, field)
                value_max = np.abs(field).max()
                fields.append((velo_max * self.N / value_max) * field_fft)
            self.set_state_spect(np.stack(fields))
        else:
            lambda0 = min(oper.lx, oper.ly) / 4.0 if length is None else length
            np.random.seed(seed)
            shape = oper.shapeK_loc
            ux_fft = np.random.random(shape) + 1j * np.random.random(shape) - 0.5 - 0.5j
            uy_fft = np.random.random(shape) + 1j * np.random.random(shape) - 0.5 - 0.5j
            ux_fft[0, 0] = 0.0
            uy_fft[0, 0] = 0.0
            oper.projection_perp(ux_fft, uy_fft)
            oper.dealiasing_variable(ux_fft)
            oper.dealiasing_variable(uy_fft)
            k0 = 2 * pi / lambda0
            ux_fft = ux_fft * H_smooth(k0 - oper.K, 1.0 * k0)
            uy_fft = uy_fft * H_smooth(k0 - oper.K, 1.0 * k0)
            ux = oper.ifft(ux_fft)
            uy = oper.ifft(uy_fft)
            vmax = np.sqrt(ux**2 + uy**2).max()
            ux = velo_max * ux / vmax
            uy = velo_max * uy / vmax
            rot_fft = oper.rotfft_from_vecfft(oper.fft(ux), oper.fft(uy))
            if self.solver == "ns2d":
                self.set_state_spect(rot_fft[None])
            else:  # init_from_rotfft: b = 0 (ns2d/strat/state.py:233-236, ns2d/bouss/state.py:140-142)
                self.set_state_spect(np.stack([rot_fft, np.zeros_like(rot_fft)]))

    def init_taylor_green(self):
        """doc/test_cases/Taylor_Green_vortices/run_simul.py:40-54."""
        oper = self.oper
        Z, Y, X = np.meshgrid(oper.z, oper.y, oper.x, indexing="ij")
        vx = np.sin(X) * np.cos(Y) * np.cos(Z)
        vy = -np.cos(X) * np.sin(Y) * np.cos(Z)
        vz = np.zeros_like(vx)
        self.set_state_spect(np.stack([oper.fft(v) for v in (vx, vy, vz)]))
