"""CPU restatement of the third-party layer under the hot path (TEST INFRASTRUCTURE).

fluidfft 0.4.2 (``/root/reference/pdm.lock:920-921``) and fluiddyn 0.6.6
(``pdm.lock:900-901``) are not vendored in /root/reference and cannot be installed
offline.  This module restates, in numpy/scipy, exactly the pieces the reference's hot
path calls, anchored on the reference call sites:

* FFT plugin class (``fft_as_arg / ifft_as_arg / fft / ifft / get_shapeK_seq / get_dimX_K
  ...``): ``fluidsim/operators/operators3d.py:120-133`` (required method list),
  ``fluidsim/solvers/ns3d/solver.py:210-241`` (aliasing rules),
  ``fluidsim/base/state.py:318-332``.
  Convention: forward r2c scaled by 1/(n0*n1*n2), inverse c2r unscaled (so that the
  k=0 mode is the spatial mean, ``fluidsim/base/state.py:385-392``).
* ``OperatorsPseudoSpectral3D`` / ``2D`` base classes: attributes consumed at
  ``fluidsim/operators/operators3d.py:207-299`` and ``operators2d.py:115-221``.
* ``vector_product`` (``solvers/ns3d/solver.py:19,226``), ``SetOfVariables``
  (``fluidsim/base/setofvariables.py:12``).

PARITY UNPINNED for this layer: no reference test stores values at this boundary.  The
comparator of the "cubic" dealiasing mask in particular is restated from memory of the
upstream source (``abs(K) >= coef * deltak * (n//2 + 1)`` per axis, OR-ed); every
pinned test therefore feeds the *mask array* of this oracle to the CUDA path.
"""

from math import pi

import numpy as np
import scipy.fft as sfft

WORKERS = -1  # all host threads


# --------------------------------------------------------------------------- fluiddyn
class SetOfVariables(np.ndarray):
    """ndarray subclass (nvar, *shape_variable) with keys (fluiddyn.calcul.setofvariables)."""

    def __new__(
        cls,
        input_array=None,
        keys=None,
        shape_variable=None,
        like=None,
        value=None,
        info=None,
        dtype=None,
        **kwargs,
    ):
        if input_array is not None:
            arr = input_array
            if keys is None:
                raise ValueError("keys required")
        elif like is not None:
            info = info if info is not None else like.info
            keys = like.keys
            shape = like.shape
            if dtype is None:
                dtype = like.dtype
            arr = np.empty(shape, dtype=dtype) if value is None else value * np.ones(shape, dtype=dtype)
        else:
            if dtype is None:
                dtype = np.float64
            shape = [len(keys)] + list(shape_variable)
            arr = np.empty(shape, dtype=dtype) if value is None else value * np.ones(shape, dtype=dtype)
        obj = np.asarray(arr).view(cls)
        obj.keys = list(keys)
        obj.nvar = len(keys)
        obj.info = info
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.keys = getattr(obj, "keys", None)
        self.nvar = getattr(obj, "nvar", None)
        self.info = getattr(obj, "info", None)

    def set_var(self, arg, value):
        index = arg if isinstance(arg, int) else self.keys.index(arg)
        self[index] = value

    def get_var(self, arg):
        index = arg if isinstance(arg, int) else self.keys.index(arg)
        return self[index].view(np.ndarray)

    def initialize(self, value=0):
        self[:] = value


# --------------------------------------------------------------------------- FFT classes
def _k_adim(n):
    """[0..n/2, -n/2+1..-1] (FFTW ordering; np.fft.fftfreq * n with +n/2 for even n)."""
    k = np.fft.fftfreq(n, 1.0 / n)
    if n % 2 == 0:
        k[n // 2] = n // 2
    return np.array(k, dtype=float)


class FFT3DNumpy:
    """Sequential 3-D real<->complex FFT object with the fluidfft plugin interface."""

    def __init__(self, n0, n1, n2):
        self.n0, self.n1, self.n2 = int(n0), int(n1), int(n2)
        self.shapeX = (self.n0, self.n1, self.n2)
        self.shapeK = (self.n0, self.n1, self.n2 // 2 + 1)
        self.coef_norm = self.n0 * self.n1 * self.n2
        self.comm = None

    # layout
    def get_short_name(self):
        return "fft3d.oracle_numpy"

    def get_shapeX_loc(self):
        return self.shapeX

    get_shapeX_seq = get_shapeX_loc

    def get_shapeK_loc(self):
        return self.shapeK

    get_shapeK_seq = get_shapeK_loc

    def get_local_size_X(self):
        return int(np.prod(self.shapeX))

    def get_local_size_K(self):
        return int(np.prod(self.shapeK))

    def get_dimX_K(self):
        return (0, 1, 2)

    def get_seq_indices_first_K(self):
        return (0, 0, 0)

    def get_seq_indices_first_X(self):
        return (0, 0, 0)

    def get_dim_first_fft(self):
        return 2

    def get_k_adim_loc(self):
        return _k_adim(self.n0), _k_adim(self.n1), np.arange(self.n2 // 2 + 1, dtype=float)

    # transforms
    def fft(self, fieldX):
        out = sfft.rfftn(fieldX, workers=WORKERS)
        out /= self.coef_norm
        return out

    def ifft(self, fieldK):
        return sfft.irfftn(fieldK, s=self.shapeX, workers=WORKERS) * self.coef_norm

    def fft_as_arg(self, fieldX, fieldK):
        fieldK[...] = self.fft(fieldX)

    def ifft_as_arg(self, fieldK, fieldX):
        fieldX[...] = self.ifft(fieldK)

    ifft_as_arg_destroy = ifft_as_arg

    # reductions
    def sum_wavenumbers(self, fieldK):
        if self.n2 % 2 == 0:
            return float(
                np.sum(fieldK[..., 0])
                + np.sum(fieldK[..., -1])
                + 2 * np.sum(fieldK[..., 1:-1])
            )
        return float(np.sum(fieldK[..., 0]) + 2 * np.sum(fieldK[..., 1:]))

    def compute_energy_from_X(self, fieldX):
        return float(np.mean(fieldX**2) / 2)

    def compute_energy_from_K(self, fieldK):
        return self.sum_wavenumbers(np.abs(fieldK) ** 2) / 2

    def create_arrayX(self, value=None, shape="loc"):
        a = np.empty(self.shapeX)
        if value is not None:
            a.fill(value)
        return a

    def create_arrayK(self, value=None, shape="loc"):
        a = np.empty(self.shapeK, dtype=np.complex128)
        if value is not None:
            a.fill(value)
        return a

    def gather_Xspace(self, a, root=None):
        return a

    def scatter_Xspace(self, a, root=None):
        return a

    def build_invariant_arrayX_from_2d_indices12X(self, o2d, arr2d):
        return np.ascontiguousarray(np.broadcast_to(arr2d, self.shapeX))

    def build_invariant_arrayK_from_2d_indices12X(self, o2d, arr2d):
        ret = np.zeros(self.shapeK, dtype=np.complex128)
        ret[0] = arr2d
        return ret


class FFT2DNumpy:
    """Sequential 2-D real<->complex FFT object (n0=ny, n1=nx), not transposed."""

    def __init__(self, n0, n1):
        self.n0, self.n1 = int(n0), int(n1)
        self.shapeX = (self.n0, self.n1)
        self.shapeK = (self.n0, self.n1 // 2 + 1)
        self.coef_norm = self.n0 * self.n1
        self.comm = None

    def get_short_name(self):
        return "fft2d.oracle_numpy"

    def get_shapeX_loc(self):
        return self.shapeX

    get_shapeX_seq = get_shapeX_loc

    def get_shapeK_loc(self):
        return self.shapeK

    get_shapeK_seq = get_shapeK_loc

    def get_is_transposed(self):
        return False

    def get_seq_indices_first_K(self):
        return (0, 0)

    def get_seq_indices_first_X(self):
        return (0, 0)

    def get_local_size_X(self):
        return int(np.prod(self.shapeX))

    def fft(self, fieldX):
        out = sfft.rfft2(fieldX, workers=WORKERS)
        out /= self.coef_norm
        return out

    def ifft(self, fieldK):
        return sfft.irfft2(fieldK, s=self.shapeX, workers=WORKERS) * self.coef_norm

    def fft_as_arg(self, fieldX, fieldK):
        fieldK[...] = self.fft(fieldX)

    def ifft_as_arg(self, fieldK, fieldX):
        fieldX[...] = self.ifft(fieldK)

    def sum_wavenumbers(self, fieldK):
        if self.n1 % 2 == 0:
            return float(
                np.sum(fieldK[..., 0])
                + np.sum(fieldK[..., -1])
                + 2 * np.sum(fieldK[..., 1:-1])
            )
        return float(np.sum(fieldK[..., 0]) + 2 * np.sum(fieldK[..., 1:]))

    def compute_energy_from_X(self, fieldX):
        return float(np.mean(fieldX**2) / 2)

    def compute_energy_from_K(self, fieldK):
        return self.sum_wavenumbers(np.abs(fieldK) ** 2) / 2


# --------------------------------------------------------------------------- 3-D operators
def vector_product(ax, ay, az, bx, by, bz):
    """a x b, written INTO bx, by, bz and returned (solver.py:226 relies on this)."""
    n0, n1, n2 = ax.shape
    elem_x = ay * bz - az * by
    elem_y = az * bx - ax * bz
    elem_z = ax * by - ay * bx
    bx[...] = elem_x
    by[...] = elem_y
    bz[...] = elem_z
    return bx, by, bz


class OperatorsPseudoSpectral3D:
    """Restated fluidfft.fft3d.operators.OperatorsPseudoSpectral3D (sequential)."""

    def __init__(self, nx, ny, nz, lx, ly, lz, fft=None, coef_dealiasing=1.0):
        self.nx = self.nx_seq = int(nx)
        self.ny = self.ny_seq = int(ny)
        self.nz = self.nz_seq = int(nz)
        self.lx = self.Lx = float(lx)
        self.ly = self.Ly = float(ly)
        self.lz = self.Lz = float(lz)

        if fft is None or isinstance(fft, str):
            op_fft = FFT3DNumpy(nz, ny, nx)
        else:
            op_fft = fft
        self._op_fft = self.oper_fft = op_fft
        self.type_fft = op_fft.__class__.__module__

        self.shapeX_seq = op_fft.get_shapeX_seq()
        self.shapeX_loc = op_fft.get_shapeX_loc()
        self.shapeK_seq = self.shapeK = op_fft.get_shapeK_seq()
        self.shapeK_loc = op_fft.get_shapeK_loc()
        self.nk0, self.nk1, self.nk2 = self.shapeK_loc

        self.deltax = self.lx / self.nx
        self.deltay = self.ly / self.ny
        self.deltaz = self.lz / self.nz
        self.x_seq = self.x = self.deltax * np.arange(self.nx)
        self.y_seq = self.y = self.deltay * np.arange(self.ny)
        self.z_seq = self.z = self.deltaz * np.arange(self.nz)

        self.deltakx = 2 * pi / self.lx
        self.deltaky = 2 * pi / self.ly
        self.deltakz = 2 * pi / self.lz

        self.ifft = self.ifft3d = op_fft.ifft
        self.fft = self.fft3d = op_fft.fft
        self.ifft_as_arg = op_fft.ifft_as_arg
        self.fft_as_arg = op_fft.fft_as_arg
        self.ifft_as_arg_destroy = getattr(op_fft, "ifft_as_arg_destroy", op_fft.ifft_as_arg)
        self.sum_wavenumbers = op_fft.sum_wavenumbers
        self.compute_energy_from_X = op_fft.compute_energy_from_X
        self.compute_energy_from_K = op_fft.compute_energy_from_K

        self.rank = 0
        self.comm = None
        self.is_sequential = True
        self._is_mpi_lib = False

        k0_adim, k1_adim, k2_adim = op_fft.get_k_adim_loc()
        self.dimX_K = op_fft.get_dimX_K()
        deltaks = (self.deltakz, self.deltaky, self.deltakx)
        self.k0 = deltaks[self.dimX_K[0]] * np.asarray(k0_adim, dtype=float)
        self.k1 = deltaks[self.dimX_K[1]] * np.asarray(k1_adim, dtype=float)
        self.k2 = deltaks[self.dimX_K[2]] * np.asarray(k2_adim, dtype=float)
        K0, K1, K2_ = np.meshgrid(self.k0, self.k1, self.k2, indexing="ij", copy=True)
        Ks = [np.ascontiguousarray(K) for K in (K0, K1, K2_)]
        assert Ks[0].shape == tuple(self.shapeK_loc)
        self.Kz = Ks[self.dimX_K.index(0)]
        self.Ky = Ks[self.dimX_K.index(1)]
        self.Kx = Ks[self.dimX_K.index(2)]

        self.K2 = self.Kx**2 + self.Ky**2 + self.Kz**2
        self.K8 = self.K2**4
        self.seq_indices_first_K = op_fft.get_seq_indices_first_K()
        self.seq_indices_first_X = op_fft.get_seq_indices_first_X()

        K_square_nozero = self.K2.copy()
        if all(i == 0 for i in self.seq_indices_first_K):
            K_square_nozero[0, 0, 0] = 1e-14
        self.inv_K_square_nozero = 1.0 / K_square_nozero
        Kh_square_nozero = self.Kx**2 + self.Ky**2
        Kh_square_nozero[Kh_square_nozero == 0] = 1e-14
        self.inv_Kh_square_nozero = 1.0 / Kh_square_nozero

        self.coef_dealiasing = coef_dealiasing
        # [EXT, unpinned] cubic truncation
        kx_max = self.deltakx * (self.nx // 2 + 1)
        ky_max = self.deltaky * (self.ny // 2 + 1)
        kz_max = self.deltakz * (self.nz // 2 + 1)
        cond = (
            (np.abs(self.Kx) >= coef_dealiasing * kx_max)
            | (np.abs(self.Ky) >= coef_dealiasing * ky_max)
            | (np.abs(self.Kz) >= coef_dealiasing * kz_max)
        )
        self.where_dealiased = np.array(cond, dtype=np.uint8)

        # spectra helpers
        self.deltak = max(self.deltakx, self.deltaky, self.deltakz)
        self.nk_spectra = int(
            np.sqrt(
                (self.deltakx * (self.nx // 2)) ** 2
                + (self.deltaky * (self.ny // 2)) ** 2
                + (self.deltakz * (self.nz // 2)) ** 2
            )
            / self.deltak
        ) + 2
        self.k_spectra3d = self.deltak * np.arange(self.nk_spectra)

    # containers
    def create_arrayX(self, value=None, shape="loc"):
        return self.oper_fft.create_arrayX(value, shape)

    def create_arrayK(self, value=None, shape="loc"):
        return self.oper_fft.create_arrayK(value, shape)

    # elementwise k-space operators (Appendix A)
    def project_perpk3d(self, vx_fft, vy_fft, vz_fft):
        tmp = (self.Kx * vx_fft + self.Ky * vy_fft + self.Kz * vz_fft) * self.inv_K_square_nozero
        vx_fft -= self.Kx * tmp
        vy_fft -= self.Ky * tmp
        vz_fft -= self.Kz * tmp

    def divfft_from_vecfft(self, vx_fft, vy_fft, vz_fft):
        return 1j * (self.Kx * vx_fft + self.Ky * vy_fft + self.Kz * vz_fft)

    def rotfft_from_vecfft(self, vx_fft, vy_fft, vz_fft):
        return (
            1j * (self.Ky * vz_fft - self.Kz * vy_fft),
            1j * (self.Kz * vx_fft - self.Kx * vz_fft),
            1j * (self.Kx * vy_fft - self.Ky * vx_fft),
        )

    def rotfft_from_vecfft_outin(self, vx_fft, vy_fft, vz_fft, rotxfft, rotyfft, rotzfft):
        rotxfft[...] = 1j * (self.Ky * vz_fft - self.Kz * vy_fft)
        rotyfft[...] = 1j * (self.Kz * vx_fft - self.Kx * vz_fft)
        rotzfft[...] = 1j * (self.Kx * vy_fft - self.Ky * vx_fft)

    def rotzfft_from_vxvyfft(self, vx_fft, vy_fft):
        return 1j * (self.Kx * vy_fft - self.Ky * vx_fft)

    def divhfft_from_vxvyfft(self, vx_fft, vy_fft):
        return 1j * (self.Kx * vx_fft + self.Ky * vy_fft)

    def div_vb_fft_from_vb(self, vx, vy, vz, b):
        fft = self.fft
        return self.divfft_from_vecfft(fft(vx * b), fft(vy * b), fft(vz * b))

    def div_vv_fft_from_v(self, vx, vy, vz):
        return (
            self.div_vb_fft_from_vb(vx, vy, vz, vx),
            self.div_vb_fft_from_vb(vx, vy, vz, vy),
            self.div_vb_fft_from_vb(vx, vy, vz, vz),
        )

    def gradfft_from_fft(self, f_fft):
        return 1j * self.Kx * f_fft, 1j * self.Ky * f_fft, 1j * self.Kz * f_fft

    # observables
    def compute_3dspectrum(self, energy_fft):
        """Shell spectrum; linear sharing between adjacent shells (Appendix A, medium)."""
        K = np.sqrt(self.K2)
        w = np.full(self.shapeK_loc, 2.0)
        w[..., 0] = 1.0
        if self.nx % 2 == 0:
            w[..., -1] = 1.0
        E = energy_fft * w
        nk = self.nk_spectra
        kappa = K / self.deltak
        ik = np.floor(kappa).astype(int)
        coef_share = kappa - ik
        spectrum = np.zeros(nk)
        last = ik >= nk - 1
        np.add.at(spectrum, np.where(last, nk - 1, ik), np.where(last, E, (1 - coef_share) * E))
        np.add.at(spectrum, np.where(last, nk - 1, ik + 1), np.where(last, 0.0, coef_share * E))
        return spectrum / self.deltak


    def compute_1dspectra(self, energy_fft):
        """E(kx), E(ky), E(kz) [EXT fluidfft, restated per SURVEY Appendix A]: r2c weights along kx,
        +-ky and +-kz folded on |k| index, each divided by its own deltak so that
        sum(E_ki) * deltaki = sum_wavenumbers(energy_fft)."""
        w = np.full(self.shapeK_loc, 2.0)
        w[..., 0] = 1.0
        if self.nx % 2 == 0:
            w[..., -1] = 1.0
        E = energy_fft * w
        nkz, nky = self.nz // 2 + 1, self.ny // 2 + 1
        e_kx = E.sum(axis=(0, 1)) / self.deltakx
        tmp_y = E.sum(axis=(0, 2))
        iy = np.rint(np.abs(self.Ky[0, :, 0]) / self.deltaky).astype(int)
        e_ky = np.zeros(nky)
        np.add.at(e_ky, iy, tmp_y)
        tmp_z = E.sum(axis=(1, 2))
        iz = np.rint(np.abs(self.Kz[:, 0, 0]) / self.deltakz).astype(int)
        e_kz = np.zeros(nkz)
        np.add.at(e_kz, iz, tmp_z)
        return e_kx, e_ky / self.deltaky, e_kz / self.deltakz


# --------------------------------------------------------------------------- 2-D operators
class OperatorsPseudoSpectral2D:
    """Restated fluidfft.fft2d.operators.OperatorsPseudoSpectral2D (sequential)."""

    def __init__(self, nx, ny, lx, ly, fft=None, coef_dealiasing=1.0):
        self.nx = self.nx_seq = int(nx)
        self.ny = self.ny_seq = int(ny)
        self.lx = float(lx)
        self.ly = float(ly)
        if fft is None or isinstance(fft, str):
            opfft = FFT2DNumpy(ny, nx)
        else:
            opfft = fft
        self.opfft = self._opfft = self.oper_fft = opfft
        self.type_fft = opfft.__class__.__module__
        self.is_transposed = opfft.get_is_transposed()
        self.is_sequential = True
        self.rank = 0
        self.shapeX = self.shapeX_seq = self.shapeX_loc = opfft.get_shapeX_loc()
        self.shapeK = self.shapeK_seq = self.shapeK_loc = opfft.get_shapeK_loc()
        self.nkx_loc = self.shapeK_loc[1]
        self.nky_loc = self.shapeK_loc[0]

        self.fft = self.fft2 = opfft.fft
        self.ifft = self.ifft2 = opfft.ifft
        self.fft_as_arg = opfft.fft_as_arg
        self.ifft_as_arg = opfft.ifft_as_arg
        self.sum_wavenumbers = opfft.sum_wavenumbers
        self.compute_energy_from_X = opfft.compute_energy_from_X
        self.compute_energy_from_K = opfft.compute_energy_from_K

        self.deltax = self.lx / self.nx
        self.deltay = self.ly / self.ny
        self.x_seq = self.x = self.x_loc = self.deltax * np.arange(self.nx)
        self.y_seq = self.y = self.y_loc = self.deltay * np.arange(self.ny)
        self.XX, self.YY = np.meshgrid(self.x, self.y)

        self.deltakx = 2 * pi / self.lx
        self.deltaky = 2 * pi / self.ly
        self.nkxE = self.nx // 2 + 1
        self.nkyE = self.ny // 2 + 1
        self.kxE = self.deltakx * np.arange(self.nkxE)
        self.kyE = self.deltaky * np.arange(self.nkyE)
        kx = self.deltakx * np.arange(self.nx // 2 + 1, dtype=float)
        ky = self.deltaky * _k_adim(self.ny)
        self.kx = self.kx_loc = kx
        self.ky = self.ky_loc = ky
        self.KX, self.KY = np.meshgrid(kx, ky)
        self.KX = np.ascontiguousarray(self.KX)
        self.KY = np.ascontiguousarray(self.KY)
        self.KX2 = self.KX**2
        self.KY2 = self.KY**2
        self.K2 = self.KX2 + self.KY2
        self.K4 = self.K2**2
        self.K8 = self.K4**2
        self.K = np.sqrt(self.K2)
        self.K2_not0 = self.K2.copy()
        self.K2_not0[0, 0] = 1e-14
        self.K_not0 = np.sqrt(self.K2_not0)
        self.K4_not0 = self.K2_not0**2
        self.inv_K2_not0 = 1.0 / self.K2_not0

        self.coef_dealiasing = coef_dealiasing
        self._has_to_dealiase = coef_dealiasing < 1.0
        # [EXT, unpinned] rectangular truncation
        kx_max = self.deltakx * (self.nx // 2 + 1)
        ky_max = self.deltaky * (self.ny // 2 + 1)
        cond = (np.abs(self.KX) >= coef_dealiasing * kx_max) | (
            np.abs(self.KY) >= coef_dealiasing * ky_max
        )
        self.where_dealiased = np.array(cond, dtype=np.uint8)
        self.deltak = max(self.deltakx, self.deltaky)

    def create_arrayX(self, value=None, shape="loc"):
        a = np.empty(self.shapeX_loc)
        if value is not None:
            a.fill(value)
        return a

    def create_arrayK(self, value=None, shape="loc"):
        a = np.empty(self.shapeK_loc, dtype=np.complex128)
        if value is not None:
            a.fill(value)
        return a

    def dealiasing_variable(self, f_fft):
        if self._has_to_dealiase:
            f_fft[np.nonzero(self.where_dealiased)] = 0.0

    def vecfft_from_rotfft(self, rot_fft):
        ux_fft = 1j * self.KY * self.inv_K2_not0 * rot_fft
        uy_fft = -1j * self.KX * self.inv_K2_not0 * rot_fft
        return ux_fft, uy_fft

    def gradfft_from_fft(self, f_fft):
        return 1j * self.KX * f_fft, 1j * self.KY * f_fft

    def rotfft_from_vecfft(self, vx_fft, vy_fft):
        return 1j * (self.KX * vy_fft - self.KY * vx_fft)

    def divfft_from_vecfft(self, vx_fft, vy_fft):
        return 1j * (self.KX * vx_fft + self.KY * vy_fft)

    def projection_perp(self, fx_fft, fy_fft):
        tmp = (self.KX * fx_fft + self.KY * fy_fft) * self.inv_K2_not0
        fx_fft -= self.KX * tmp
        fy_fft -= self.KY * tmp
        return fx_fft, fy_fft

    def laplacian_fft(self, a_fft, order=2):
        return (-1) ** (order // 2) * self.K2 ** (order // 2) * a_fft
