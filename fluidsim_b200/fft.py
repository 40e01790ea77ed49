"""fluidfft-plugin-shaped FFT classes backed by libb200spectral (``type_fft = "fft3d.with_b200"``).

Mirror of the plugin interface fluidsim consumes
(``/root/reference/fluidsim/operators/operators3d.py:120-133``; call sites
``solvers/ns3d/solver.py:210-241``, ``base/state.py:318-332``): ``fft_as_arg``, ``ifft_as_arg``,
``ifft_as_arg_destroy``, ``fft``, ``ifft``, ``get_shapeX_loc/seq``, ``get_shapeK_loc/seq``,
``get_dimX_K``, ``get_seq_indices_first_K/X``, ``get_k_adim_loc``, ``sum_wavenumbers``,
``compute_energy_from_X/K``, ``create_arrayX/K`` ...

Arrays may be CUDA ``torch`` tensors (device-resident, zero copy) or host ``numpy`` arrays (the
reference's calling convention: the transform then includes the host<->device copies).
Sequential layout: X ``(n0, n1, n2)`` float64, K ``(n0, n1, n2//2+1)`` complex128,
``dimX_K = (0, 1, 2)``; forward scaled by ``1/(n0 n1 n2)``, inverse unscaled.
"""

import ctypes as C
from math import pi

import numpy as np
import torch

from . import _lib
from ._lib import call, lib, ptr, stream_ptr


def _k_adim(n):
    k = np.fft.fftfreq(n, 1.0 / n)
    if n % 2 == 0:
        k[n // 2] = n // 2
    return k


class Plan:
    """Owner of a ``b2_plan*``."""

    def __init__(self, ndim, shapeX, lengths, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("fluidsim_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.ndim = ndim
        n = list(shapeX) + [0] * (3 - ndim)
        L = list(lengths) + [0.0] * (3 - ndim)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            call("b2_plan_create", C.byref(handle), ndim, n[0], n[1], n[2], L[0], L[1], L[2])
        self.handle = handle
        self.shapeX = tuple(int(x) for x in shapeX)
        self.shapeK = tuple(self.shapeX[:-1]) + (self.shapeX[-1] // 2 + 1,)
        self.is_fast = bool(lib.b2_plan_is_fast(handle))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                lib.b2_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class _FFTWithB200Base:
    ndim = None

    def _init(self, shapeX, lengths):
        self.plan = Plan(self.ndim, shapeX, lengths)
        self.device = self.plan.device
        self.shapeX = self.plan.shapeX
        self.shapeK = self.plan.shapeK
        self.coef_norm = int(np.prod(self.shapeX))
        self._work = None
        self._stageX = None
        self._stageK = None
        self._scalar = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.comm = None

    # ---- layout ------------------------------------------------------------------------------
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

    def get_seq_indices_first_K(self):
        return (0,) * self.ndim

    def get_seq_indices_first_X(self):
        return (0,) * self.ndim

    def create_arrayX(self, value=None, shape="loc"):
        a = torch.empty(self.shapeX, dtype=torch.float64, device=self.device)
        if value is not None:
            a.fill_(value)
        return a

    def create_arrayK(self, value=None, shape="loc"):
        a = torch.empty(self.shapeK, dtype=torch.complex128, device=self.device)
        if value is not None:
            a.fill_(value)
        return a

    # ---- staging for host (numpy) callers ------------------------------------------------------
    def _devX(self, x):
        if isinstance(x, torch.Tensor):
            if x.dtype != torch.float64 or not x.is_cuda:
                raise TypeError("fieldX must be a CUDA float64 tensor")
            return x
        if self._stageX is None:
            self._stageX = self.create_arrayX()
        self._stageX.copy_(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)))
        return self._stageX

    def _devK(self, k):
        if isinstance(k, torch.Tensor):
            if k.dtype != torch.complex128 or not k.is_cuda:
                raise TypeError("fieldK must be a CUDA complex128 tensor")
            return k
        if self._stageK is None:
            self._stageK = self.create_arrayK()
        self._stageK.copy_(torch.from_numpy(np.ascontiguousarray(k, dtype=np.complex128)))
        return self._stageK

    def _check_shapes(self, fieldX=None, fieldK=None):
        if fieldX is not None and tuple(fieldX.shape) != self.shapeX:
            raise ValueError(f"fieldX has shape {tuple(fieldX.shape)}, expected {self.shapeX}")
        if fieldK is not None and tuple(fieldK.shape) != self.shapeK:
            raise ValueError(f"fieldK has shape {tuple(fieldK.shape)}, expected {self.shapeK}")

    # ---- transforms ------------------------------------------------------------------------------
    def fft_as_arg(self, fieldX, fieldK):
        """fieldK <- FFT(fieldX) / N  (caller-owned output, written in place)."""
        self._check_shapes(fieldX, fieldK)
        xd = self._devX(fieldX)
        host_out = not isinstance(fieldK, torch.Tensor)
        if host_out:
            if self._stageK is None:
                self._stageK = self.create_arrayK()
            kd = self._stageK
        else:
            kd = self._devK(fieldK)
        call("b2_fft_r2c", self.plan.handle, ptr(xd), ptr(kd), stream_ptr())
        if host_out:
            fieldK[...] = kd.cpu().numpy()

    def ifft_as_arg(self, fieldK, fieldX):
        """fieldX <- unnormalised inverse FFT of fieldK; fieldK is left intact."""
        self._check_shapes(fieldX, fieldK)
        kd = self._devK(fieldK)
        if self._work is None:
            self._work = self.create_arrayK()
        self._c2r(kd, fieldX, self._work)

    def ifft_as_arg_destroy(self, fieldK, fieldX):
        """Same, but fieldK may be clobbered (used on scratch arrays, solver.py:210-212)."""
        self._check_shapes(fieldX, fieldK)
        kd = self._devK(fieldK)
        self._c2r(kd, fieldX, None)

    def _c2r(self, kd, fieldX, work):
        host_out = not isinstance(fieldX, torch.Tensor)
        if host_out:
            if self._stageX is None:
                self._stageX = self.create_arrayX()
            xd = self._stageX
        else:
            xd = self._devX(fieldX)
        call("b2_ifft_c2r", self.plan.handle, ptr(kd), ptr(xd), ptr(work), stream_ptr())
        if host_out:
            fieldX[...] = xd.cpu().numpy()

    def fft(self, fieldX):
        if isinstance(fieldX, torch.Tensor):
            out = self.create_arrayK()
        else:
            out = np.empty(self.shapeK, dtype=np.complex128)
        self.fft_as_arg(fieldX, out)
        return out

    def ifft(self, fieldK):
        if isinstance(fieldK, torch.Tensor):
            out = self.create_arrayX()
        else:
            out = np.empty(self.shapeX, dtype=np.float64)
        self.ifft_as_arg(fieldK, out)
        return out

    # ---- reductions --------------------------------------------------------------------------------
    def sum_wavenumbers(self, fieldK):
        """r2c-aware sum over wavenumbers of a REAL K-shaped array (fluidfft semantics)."""
        if isinstance(fieldK, torch.Tensor):
            a = fieldK
            n_last = self.shapeX[-1]
            if n_last % 2 == 0:
                s = a[..., 0].sum() + a[..., -1].sum() + 2 * a[..., 1:-1].sum()
            else:
                s = a[..., 0].sum() + 2 * a[..., 1:].sum()
            return float(s)
        a = np.asarray(fieldK)
        if self.shapeX[-1] % 2 == 0:
            return float(a[..., 0].sum() + a[..., -1].sum() + 2 * a[..., 1:-1].sum())
        return float(a[..., 0].sum() + 2 * a[..., 1:].sum())

    def sum_wavenumbers_abs2(self, fieldsK):
        """sum_wavenumbers(|f|^2) summed over the leading axis (CUDA reduction kernel)."""
        kd = fieldsK if isinstance(fieldsK, torch.Tensor) else self._devK(fieldsK)
        nvar = 1 if kd.dim() == self.ndim else kd.shape[0]
        call("b2_sum_wavenumbers_abs2", self.plan.handle, ptr(kd), nvar, ptr(self._scalar), stream_ptr())
        return float(self._scalar.item())

    def compute_energy_from_K(self, fieldK):
        return 0.5 * self.sum_wavenumbers_abs2(fieldK)

    def compute_energy_from_X(self, fieldX):
        xd = self._devX(fieldX)
        return float((xd * xd).mean().item() / 2)

    def gather_Xspace(self, a, root=None):
        return a

    def scatter_Xspace(self, a, root=None):
        return a

    def run_tests(self):
        x = torch.rand(self.shapeX, dtype=torch.float64, device=self.device)
        k = self.fft(x)
        x2 = self.ifft(k)
        err = float((x - x2).abs().max())
        if err > 1e-12:
            raise RuntimeError(f"fft/ifft round trip error {err}")
        return 0


class FFT3DWithB200(_FFTWithB200Base):
    """``FFTclass(n0, n1, n2)`` -- n0 = nz, n1 = ny, n2 = nx."""

    ndim = 3

    def __init__(self, n0, n1, n2, lengths=(2 * pi, 2 * pi, 2 * pi)):
        self.n0, self.n1, self.n2 = int(n0), int(n1), int(n2)
        self._init((self.n0, self.n1, self.n2), lengths)

    def get_short_name(self):
        return "fft3d.with_b200"

    def get_dimX_K(self):
        return (0, 1, 2)

    def get_dim_first_fft(self):
        return 2

    def get_k_adim_loc(self):
        return _k_adim(self.n0), _k_adim(self.n1), np.arange(self.n2 // 2 + 1, dtype=float)

    def build_invariant_arrayX_from_2d_indices12X(self, o2d, arr2d):
        a = arr2d if isinstance(arr2d, torch.Tensor) else torch.from_numpy(np.asarray(arr2d)).to(self.device)
        return a.unsqueeze(0).expand(self.shapeX).contiguous()

    def build_invariant_arrayK_from_2d_indices12X(self, o2d, arr2d):
        a = arr2d if isinstance(arr2d, torch.Tensor) else torch.from_numpy(np.asarray(arr2d)).to(self.device)
        ret = self.create_arrayK(0)
        ret[0] = a
        return ret


class FFT2DWithB200(_FFTWithB200Base):
    """``FFTclass(n0, n1)`` -- n0 = ny, n1 = nx; not transposed."""

    ndim = 2

    def __init__(self, n0, n1, lengths=(2 * pi, 2 * pi)):
        self.n0, self.n1 = int(n0), int(n1)
        self._init((self.n0, self.n1), lengths)

    def get_short_name(self):
        return "fft2d.with_b200"

    def get_is_transposed(self):
        return False

    def get_k_adim_loc(self):
        return _k_adim(self.n0), np.arange(self.n1 // 2 + 1, dtype=float)


# the attribute fluidfft's plugin loader looks for (``fluidfft.import_fft_class``)
FFTclass = FFT3DWithB200
