"""GPU-resident ``OperatorsPseudoSpectral3D`` / ``OperatorsPseudoSpectral2D``.

Host-side mirror of ``/root/reference/fluidsim/operators/operators3d.py:116-342`` and
``operators2d.py:85-221`` (which subclass fluidfft's operator classes): same constructor
(``params``), same attribute and method names, same error behaviour, but every array is a CUDA
``torch`` tensor and every elementwise operator is a kernel of ``libb200spectral`` called through
the C ABI.  Field-sized coefficient arrays (``Kx, Ky, Kz, K2, K4, K8, inv_K_square_nozero``) are
lazy properties: the kernels recompute wavenumbers from three 1-D vectors instead of loading 3-D
coefficient arrays (SURVEY.md section 7 "hard parts").
"""

from math import pi

import numpy as np
import torch

from ._lib import call, ptr, stream_ptr
from .fft import FFT2DWithB200, FFT3DWithB200, _k_adim
from .setofvariables import SetOfVariables


def vector_product(ax, ay, az, bx, by, bz):
    """a x b written INTO bx, by, bz (fluidfft semantics, solvers/ns3d/solver.py:226)."""
    call("b2_vector_product", ptr(ax), ptr(ay), ptr(az), ptr(bx), ptr(by), ptr(bz), ax.numel(), stream_ptr())
    return bx, by, bz


def _check_type_fft(type_fft, ndim):
    ok = (None, "default", "sequential", f"fft{ndim}d.with_b200", f"fluidsim_b200.fft{ndim}d.with_b200")
    if type_fft not in ok:
        raise ValueError(
            f"type_fft = {type_fft!r}: fluidsim_b200 only provides 'fft{ndim}d.with_b200' (no CPU fallback)"
        )


class _OperatorBase:
    """``OperatorBase._reinit_truncation`` / ``mean_space`` (operators/base.py:47-81)."""

    def _reinit_truncation(self):
        try:
            truncation_shape = self.params.oper.truncation_shape
        except AttributeError:
            return
        if truncation_shape == "cubic":
            return
        kmax = self.coef_dealiasing * self.deltakx * self.nx / 2
        if truncation_shape == "spherical":
            self.where_dealiased = (self.K2 >= kmax**2).to(torch.uint8)
        elif truncation_shape == "no_multiple_aliases":
            where = self.get_region_multiple_aliases()
            if self.coef_dealiasing:
                where |= self.K2 >= kmax**2
            self.where_dealiased = where.to(torch.uint8)
        else:
            raise ValueError(
                'truncation_shape must be "cubic", "spherical" or "no_multiple_aliases"'
            )

    def mean_space(self, arr):
        return float(arr.mean())


class OperatorsPseudoSpectral3D(_OperatorBase):
    def __init__(self, params=None):
        self.params = params
        self.axes = ("z", "y", "x")
        po = params.oper
        po.nx, po.ny, po.nz = int(po.nx), int(po.ny), int(po.nz)
        if getattr(params, "ONLY_COARSE_OPER", False):
            nx = ny = nz = 4
        else:
            nx, ny, nz = po.nx, po.ny, po.nz
        _check_type_fft(getattr(po, "type_fft", "default"), 3)
        self.nx = self.nx_seq = nx
        self.ny = self.ny_seq = ny
        self.nz = self.nz_seq = nz
        self.Lx = self.lx = float(po.Lx)
        self.Ly = self.ly = float(po.Ly)
        self.Lz = self.lz = float(po.Lz)
        self.oper_fft = self._op_fft = FFT3DWithB200(nz, ny, nx, lengths=(self.Lz, self.Ly, self.Lx))
        op = self.oper_fft
        self.plan = op.plan
        self.device = op.device
        self.type_fft = "fluidsim_b200.fft3d.with_b200"
        self.shapeX = self.shapeX_seq = self.shapeX_loc = op.get_shapeX_loc()
        self.shapeK = self.shapeK_seq = self.shapeK_loc = op.get_shapeK_loc()
        self.nk0, self.nk1, self.nk2 = self.shapeK_loc
        self.dimX_K = op.get_dimX_K()
        self.seq_indices_first_K = op.get_seq_indices_first_K()
        self.seq_indices_first_X = op.get_seq_indices_first_X()
        self.is_sequential = True
        self._is_mpi_lib = False
        self.rank = 0
        self.comm = None
        self.SAME_SIZE_IN_ALL_PROC = True

        self.deltax, self.deltay, self.deltaz = self.Lx / nx, self.Ly / ny, self.Lz / nz
        self.x_seq = self.x = self.deltax * np.arange(nx)
        self.y_seq = self.y = self.deltay * np.arange(ny)
        self.z_seq = self.z = self.deltaz * np.arange(nz)
        self.deltakx, self.deltaky, self.deltakz = 2 * pi / self.Lx, 2 * pi / self.Ly, 2 * pi / self.Lz
        self.deltak = max(self.deltakx, self.deltaky, self.deltakz)
        # 1-D dimensional wavenumbers along K axes 0, 1, 2
        self.k0 = self.deltakz * _k_adim(nz)
        self.k1 = self.deltaky * _k_adim(ny)
        self.k2 = self.deltakx * np.arange(nx // 2 + 1, dtype=float)
        dev = self.device
        self._k0d = torch.from_numpy(self.k0).to(dev)
        self._k1d = torch.from_numpy(self.k1).to(dev)
        self._k2d = torch.from_numpy(self.k2).to(dev)

        # transforms
        self.fft = self.fft3d = op.fft
        self.ifft = self.ifft3d = op.ifft
        self.fft_as_arg = op.fft_as_arg
        self.ifft_as_arg = op.ifft_as_arg
        self.ifft_as_arg_destroy = op.ifft_as_arg_destroy
        self.sum_wavenumbers = op.sum_wavenumbers
        self.compute_energy_from_X = op.compute_energy_from_X
        self.compute_energy_from_K = op.compute_energy_from_K
        self.create_arrayX = op.create_arrayX
        self.create_arrayK = op.create_arrayK

        self.coef_dealiasing = po.coef_dealiasing
        self.where_dealiased = self._cubic_mask()
        self._reinit_truncation()
        if getattr(po, "NO_SHEAR_MODES", False):
            cond = (self._k2d[None, None, :] ** 2 + self._k1d[None, :, None] ** 2) == 0.0
            self.where_dealiased = (cond | self.where_dealiased.bool()).to(torch.uint8).contiguous()
        self._tmpK = None

    # ---- coefficient arrays (lazy) -----------------------------------------------------------------
    def _cubic_mask(self):
        """fluidfft's default ("cubic") truncation.  [EXT, unpinned]: restated as
        ``abs(K_i) >= coef * deltak_i * (n_i // 2 + 1)`` OR-ed over the three axes; the fused kernels
        take this mask as an input array, so a different upstream definition only requires
        assigning ``oper.where_dealiased``."""
        c = self.coef_dealiasing
        cx = self._k2d.abs() >= c * self.deltakx * (self.nx // 2 + 1)
        cy = self._k1d.abs() >= c * self.deltaky * (self.ny // 2 + 1)
        cz = self._k0d.abs() >= c * self.deltakz * (self.nz // 2 + 1)
        m = cz[:, None, None] | cy[None, :, None] | cx[None, None, :]
        return m.to(torch.uint8).contiguous()

    @property
    def Kx(self):
        return self._k2d[None, None, :].expand(self.shapeK_loc).contiguous()

    @property
    def Ky(self):
        return self._k1d[None, :, None].expand(self.shapeK_loc).contiguous()

    @property
    def Kz(self):
        return self._k0d[:, None, None].expand(self.shapeK_loc).contiguous()

    @property
    def K2(self):
        return (
            self._k2d[None, None, :] ** 2 + self._k1d[None, :, None] ** 2 + self._k0d[:, None, None] ** 2
        ).contiguous()

    @property
    def K4(self):
        return self.K2**2

    @property
    def K8(self):
        return self.K2**4

    @property
    def K2_not0(self):
        K2 = self.K2
        K2[0, 0, 0] = 1e-14
        return K2

    @property
    def inv_K_square_nozero(self):
        return 1.0 / self.K2_not0

    def get_region_multiple_aliases(self):
        ax = self._k2d.abs() >= 2 / 3 * self.deltakx * self.nx / 2
        ay = self._k1d.abs() >= 2 / 3 * self.deltaky * self.ny / 2
        az = self._k0d.abs() >= 2 / 3 * self.deltakz * self.nz / 2
        ax, ay, az = ax[None, None, :], ay[None, :, None], az[:, None, None]
        return (ax & ay) | (ay & az) | (az & ax)

    # ---- k-space operators ---------------------------------------------------------------------------
    def project_perpk3d(self, vx_fft, vy_fft, vz_fft):
        call("b2_project_perpk3d", self.plan.handle, ptr(vx_fft), ptr(vy_fft), ptr(vz_fft), stream_ptr())

    def project_toroidal(self, vx_fft, vy_fft, vz_fft):
        """operators3d.py:911-958 (in place)."""
        call("b2_project_toroidal", self.plan.handle, ptr(vx_fft), ptr(vy_fft), ptr(vz_fft), stream_ptr())

    def project_poloidal(self, vx_fft, vy_fft, vz_fft):
        """operators3d.py:788-856 (in place)."""
        call("b2_project_poloidal", self.plan.handle, ptr(vx_fft), ptr(vy_fft), ptr(vz_fft), stream_ptr())

    def rotfft_from_vecfft_outin(self, vx_fft, vy_fft, vz_fft, rotxfft, rotyfft, rotzfft):
        call(
            "b2_rotfft_from_vecfft", self.plan.handle, ptr(vx_fft), ptr(vy_fft), ptr(vz_fft),
            ptr(rotxfft), ptr(rotyfft), ptr(rotzfft), stream_ptr(),
        )

    def rotfft_from_vecfft(self, vx_fft, vy_fft, vz_fft):
        out = tuple(self.create_arrayK() for _ in range(3))
        self.rotfft_from_vecfft_outin(vx_fft, vy_fft, vz_fft, *out)
        return out

    def divfft_from_vecfft(self, vx_fft, vy_fft, vz_fft):
        out = self.create_arrayK()
        call("b2_divfft_from_vecfft", self.plan.handle, ptr(vx_fft), ptr(vy_fft), ptr(vz_fft), ptr(out), stream_ptr())
        return out

    def div_vb_fft_from_vb(self, vx, vy, vz, b):
        """divfft_from_vecfft(fft(vx b), fft(vy b), fft(vz b)) (strat/solver.py:206)."""
        prods = []
        tmp = self.create_arrayX()
        for v in (vx, vy, vz):
            call("b2_mul_real", ptr(v), ptr(b), ptr(tmp), tmp.numel(), stream_ptr())
            prods.append(self.fft(tmp))
        return self.divfft_from_vecfft(*prods)

    def dealiasing(self, *args):
        """operators3d.py:336-342."""
        for thing in args:
            if isinstance(thing, SetOfVariables):
                call("b2_dealias", self.plan.handle, ptr(thing.tensor), thing.nvar, ptr(self.where_dealiased), stream_ptr())
            elif isinstance(thing, torch.Tensor):
                nvar = 1 if thing.dim() == 3 else thing.shape[0]
                call("b2_dealias", self.plan.handle, ptr(thing), nvar, ptr(self.where_dealiased), stream_ptr())

    # ---- observables -----------------------------------------------------------------------------------
    def compute_energy_from_3fields(self, vx_fft, vy_fft, vz_fft):
        return 0.5 * (vx_fft.abs() ** 2 + vy_fft.abs() ** 2 + vz_fft.abs() ** 2)

    def compute_3dspectrum(self, energy_fft):
        """Shell spectrum with linear sharing between adjacent shells (fluidfft semantics, SURVEY
        Appendix A); small reduction done with torch on the device."""
        K = torch.sqrt(self.K2)
        w = torch.full(self.shapeK_loc, 2.0, dtype=torch.float64, device=self.device)
        w[..., 0] = 1.0
        if self.nx % 2 == 0:
            w[..., -1] = 1.0
        E = (energy_fft * w).reshape(-1)
        nk = self.nk_spectra
        kappa = (K / self.deltak).reshape(-1)
        ik = torch.floor(kappa).long()
        share = kappa - ik
        last = ik >= nk - 1
        spectrum = torch.zeros(nk, dtype=torch.float64, device=self.device)
        spectrum.index_add_(0, torch.where(last, nk - 1, ik), torch.where(last, E, (1 - share) * E))
        spectrum.index_add_(0, torch.where(last, nk - 1, ik + 1), torch.where(last, torch.zeros_like(E), share * E))
        return (spectrum / self.deltak).cpu().numpy()

    def compute_observables(self, fields, nvar=None):
        """Everything the periodic outputs reduce from ``fields`` (a ``(nvar, *shapeK)`` tensor) in ONE
        kernel pass (C ABI ``b2_observables``): component energies, dissipation rates, enstrophy,
        per-component 3-D shell spectra and 1-D spectra.  Returns a dict of floats / numpy arrays with
        the names of the reference's outputs (solvers/ns3d/output/spatial_means.py:23-73,
        output/spectra.py:15-60).  The viscosities are the ones last pushed with b2_set_physics."""
        import ctypes as C

        from ._lib import lib

        t = fields.tensor if hasattr(fields, "tensor") else fields
        nvar = int(t.shape[0]) if nvar is None else int(nvar)
        nks = self.nk_spectra
        h = self.plan.handle
        n = int(lib.b2_observables_size(h, nvar, nks))
        out = torch.empty(n, dtype=torch.float64, device=self.device)
        call("b2_observables", h, ptr(t), nvar, nks, float(self.deltak), ptr(out), stream_ptr())
        o = out.cpu().numpy()
        names = ["vx", "vy", "vz", "b"][:nvar]
        res = {"E" + ("xyz"[i] if i < 3 else "b"): float(o[i]) for i in range(nvar)}
        res["E"] = float(sum(o[: min(nvar, 3)]))
        res.update(epsK=float(o[4]), epsK_hypo=float(o[5]), epsK4=float(o[6]), epsK8=float(o[7]),
                   enstrophy=float(o[8]))
        nkx1, nky1, nkz1 = self.nx // 2 + 1, self.ny // 2 + 1, self.nz // 2 + 1
        pos = 16
        for nm, ln in (("", nks), ("_kx", nkx1), ("_ky", nky1), ("_kz", nkz1)):
            for v in range(nvar):
                res[names[v] + nm] = o[pos:pos + ln].copy()
                pos += ln
        res["E_spectrum3d"] = sum(res[k] for k in names[:3])
        for ax in ("kx", "ky", "kz"):
            res["E_" + ax] = sum(res[k + "_" + ax] for k in names[:3])
        return res

    def compute_1dspectra(self, energy_fft):
        """fluidfft compute_1dspectra [EXT]: from an energy density array (kept for API parity; the
        fused reduction above works on the fields themselves)."""
        w = torch.full(self.shapeK_loc, 2.0, dtype=torch.float64, device=self.device)
        w[..., 0] = 1.0
        if self.nx % 2 == 0:
            w[..., -1] = 1.0
        E = energy_fft * w
        e_kx = E.sum(dim=(0, 1)) / self.deltakx
        iy = torch.round(self._k1d.abs() / self.deltaky).long()
        iz = torch.round(self._k0d.abs() / self.deltakz).long()
        e_ky = torch.zeros(self.ny // 2 + 1, dtype=torch.float64, device=self.device).index_add_(0, iy, E.sum(dim=(0, 2)))
        e_kz = torch.zeros(self.nz // 2 + 1, dtype=torch.float64, device=self.device).index_add_(0, iz, E.sum(dim=(1, 2)))
        return e_kx.cpu().numpy(), (e_ky / self.deltaky).cpu().numpy(), (e_kz / self.deltakz).cpu().numpy()

    @property
    def nk_spectra(self):
        return (
            int(
                np.sqrt(
                    (self.deltakx * (self.nx // 2)) ** 2
                    + (self.deltaky * (self.ny // 2)) ** 2
                    + (self.deltakz * (self.nz // 2)) ** 2
                )
                / self.deltak
            )
            + 2
        )


class OperatorsPseudoSpectral2D(_OperatorBase):
    def __init__(self, params):
        self.params = params
        self.axes = ("y", "x")
        po = params.oper
        nx, ny = int(po.nx), int(po.ny)
        if po.nx != nx:
            raise ValueError(f"params.oper.nx != int(params.oper.nx); ({po.nx})")
        if po.ny != ny:
            raise ValueError(f"params.oper.ny != int(params.oper.ny); ({po.ny})")
        po.nx, po.ny = nx, ny
        if getattr(params, "ONLY_COARSE_OPER", False):
            nx = ny = 4
        _check_type_fft(getattr(po, "type_fft", "default"), 2)
        self.nx = self.nx_seq = nx
        self.ny = self.ny_seq = ny
        self.Lx = self.lx = float(po.Lx)
        self.Ly = self.ly = float(po.Ly)
        self.oper_fft = self.opfft = self._opfft = FFT2DWithB200(ny, nx, lengths=(self.Ly, self.Lx))
        op = self.oper_fft
        self.plan = op.plan
        self.device = op.device
        self.type_fft = "fluidsim_b200.fft2d.with_b200"
        self.is_transposed = False
        self.is_sequential = True
        self.rank = 0
        self.shapeX = self.shapeX_seq = self.shapeX_loc = op.get_shapeX_loc()
        self.shapeK = self.shapeK_seq = self.shapeK_loc = op.get_shapeK_loc()
        self.nky_loc, self.nkx_loc = self.shapeK_loc
        self.deltax, self.deltay = self.Lx / nx, self.Ly / ny
        self.x_seq = self.x = self.deltax * np.arange(nx)
        self.y_seq = self.y = self.deltay * np.arange(ny)
        self.deltakx, self.deltaky = 2 * pi / self.Lx, 2 * pi / self.Ly
        self.deltak = max(self.deltakx, self.deltaky)
        self.kx = self.kx_loc = self.deltakx * np.arange(nx // 2 + 1, dtype=float)
        self.ky = self.ky_loc = self.deltaky * _k_adim(ny)
        dev = self.device
        self._kxd = torch.from_numpy(self.kx).to(dev)
        self._kyd = torch.from_numpy(self.ky).to(dev)

        self.fft = self.fft2 = op.fft
        self.ifft = self.ifft2 = op.ifft
        self.fft_as_arg = op.fft_as_arg
        self.ifft_as_arg = op.ifft_as_arg
        self.sum_wavenumbers = op.sum_wavenumbers
        self.compute_energy_from_X = op.compute_energy_from_X
        self.compute_energy_from_K = op.compute_energy_from_K
        self.create_arrayX = op.create_arrayX
        self.create_arrayK = op.create_arrayK

        self.coef_dealiasing = po.coef_dealiasing
        self._has_to_dealiase = self.coef_dealiasing < 1.0
        c = self.coef_dealiasing
        cx = self._kxd.abs() >= c * self.deltakx * (nx // 2 + 1)
        cy = self._kyd.abs() >= c * self.deltaky * (ny // 2 + 1)
        self.where_dealiased = (cy[:, None] | cx[None, :]).to(torch.uint8).contiguous()
        self._reinit_truncation()
        if getattr(po, "NO_SHEAR_MODES", False):
            cond = (self._kxd.abs() == 0.0)[None, :].expand(self.shapeK_loc)
            self.where_dealiased = (cond | self.where_dealiased.bool()).to(torch.uint8).contiguous()
        if getattr(po, "NO_KY0", False):
            cond = (self._kyd.abs() == 0.0)[:, None].expand(self.shapeK_loc)
            self.where_dealiased = (cond | self.where_dealiased.bool()).to(torch.uint8).contiguous()

    @property
    def KX(self):
        return self._kxd[None, :].expand(self.shapeK_loc).contiguous()

    @property
    def KY(self):
        return self._kyd[:, None].expand(self.shapeK_loc).contiguous()

    @property
    def K2(self):
        return (self._kxd[None, :] ** 2 + self._kyd[:, None] ** 2).contiguous()

    @property
    def K4(self):
        return self.K2**2

    @property
    def K8(self):
        return self.K4**2

    @property
    def K2_not0(self):
        K2 = self.K2
        K2[0, 0] = 1e-14
        return K2

    @property
    def K(self):
        return torch.sqrt(self.K2)

    def get_region_multiple_aliases(self):
        ax = (self._kxd.abs() >= 2 / 3 * self.deltakx * self.nx / 2)[None, :]
        ay = (self._kyd.abs() >= 2 / 3 * self.deltaky * self.ny / 2)[:, None]
        return ax & ay

    def vecfft_from_rotfft(self, rot_fft):
        ux, uy = self.create_arrayK(), self.create_arrayK()
        call("b2_vecfft_from_rotfft2d", self.plan.handle, ptr(rot_fft), ptr(ux), ptr(uy), stream_ptr())
        return ux, uy

    def gradfft_from_fft(self, f_fft):
        px, py = self.create_arrayK(), self.create_arrayK()
        call("b2_gradfft_from_fft2d", self.plan.handle, ptr(f_fft), ptr(px), ptr(py), stream_ptr())
        return px, py

    def rotfft_from_vecfft(self, ux_fft, uy_fft):
        rot = self.create_arrayK()
        call("b2_rotfft_from_vecfft2d", self.plan.handle, ptr(ux_fft), ptr(uy_fft), ptr(rot), stream_ptr())
        return rot

    def dealiasing(self, *args):
        """operators2d.py:200-220."""
        if not self._has_to_dealiase:
            return
        for thing in args:
            if isinstance(thing, SetOfVariables):
                call("b2_dealias", self.plan.handle, ptr(thing.tensor), thing.nvar, ptr(self.where_dealiased), stream_ptr())
            elif isinstance(thing, torch.Tensor):
                nvar = 1 if thing.dim() == 2 else thing.shape[0]
                call("b2_dealias", self.plan.handle, ptr(thing), nvar, ptr(self.where_dealiased), stream_ptr())

    def dealiasing_variable(self, f_fft):
        self.dealiasing(f_fft)
