"""ctypes binding of ``libb200spectral.so`` (the C ABI declared in ``include/b200spectral.h``).

There is NO fallback: if the shared library is missing, importing this module raises.  Build it
with ``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C fluidsim_b200/csrc``.
"""

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200spectral.so")


class B200SpectralError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: the CUDA extension must be built first "
        "(make -C fluidsim_b200/csrc).  fluidsim_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

_p = C.c_void_p
_i = C.c_int
_d = C.c_double
_ll = C.c_longlong

# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "b2_last_error": [],
    "b2_version": [],
    "b2_launch_count": [],
    "b2_plan_create": [C.POINTER(_p), _i, _i, _i, _i, _d, _d, _d],
    "b2_plan_destroy": [_p],
    "b2_plan_shapes": [_p, C.POINTER(_i), C.POINTER(_i)],
    "b2_plan_is_fast": [_p],
    "b2_fft_r2c": [_p, _p, _p, _p],
    "b2_ifft_c2r": [_p, _p, _p, _p, _p],
    "b2_rotfft_from_vecfft": [_p, _p, _p, _p, _p, _p, _p, _p],
    "b2_divfft_from_vecfft": [_p, _p, _p, _p, _p, _p],
    "b2_project_perpk3d": [_p, _p, _p, _p, _p],
    "b2_vector_product": [_p, _p, _p, _p, _p, _p, _ll, _p],
    "b2_mul_real": [_p, _p, _p, _ll, _p],
    "b2_dealias": [_p, _p, _i, _p, _p],
    "b2_vecfft_from_rotfft2d": [_p, _p, _p, _p, _p],
    "b2_gradfft_from_fft2d": [_p, _p, _p, _p, _p],
    "b2_rotfft_from_vecfft2d": [_p, _p, _p, _p, _p],
    "b2_compute_frot": [_p, _p, _p, _p, _d, _p, _ll, _p],
    "b2_tendencies_ns2d_buoyancy": [_p, _p, _p, _p, _p, _p, _d, _i, _p, _p, _ll, _p],
    "b2_compute_fb_fft": [_p, _d, _p, _ll, _p],
    "b2_add_inplace": [_p, _p, _ll, _p],
    "b2_exact_coefs": [_p, _d, _d, _d, _d, _d, _p, _p, _p],
    "b2_step_euler": [_p, _p, _d, _p, _p, _p, _i, _p],
    "b2_step_like_rk2": [_p, _p, _d, _p, _p, _p, _i, _p],
    "b2_rk4_step1": [_p, _p, _p, _p, _p, _p, _d, _i, _p],
    "b2_rk4_step2": [_p, _p, _p, _p, _p, _p, _p, _d, _i, _p],
    "b2_rk4_step3": [_p, _p, _p, _p, _d, _i, _p],
    "b2_sum_wavenumbers_abs2": [_p, _p, _i, _p, _p],
    "b2_observables_size": [_p, _i, _i],
    "b2_observables": [_p, _p, _i, _i, _d, _p, _p],
    "b2_max_abs": [_p, _ll, _p, _p],
    "b2_sum": [_p, _ll, _p, _p],
    "b2_set_physics": [_p, _i, _d, _d, _d, _d, _i, _d, _d, _d, _p],
    "b2_set_pruning": [_p, _i],
    "b2_set_no_vz_kz0": [_p, _i],
    "b2_set_projection": [_p, _i],
    "b2_project_toroidal": [_p, _p, _p, _p, _p],
    "b2_project_poloidal": [_p, _p, _p, _p, _p],
    "b2_get_pruning_bounds": [_p, C.POINTER(_i)],
    "b2_check_dealiased": [_p, _p, _i, _p, _p, _p],
    "b2_work_fields": [_p, _i, C.POINTER(_i), C.POINTER(_i)],
    "b2_set_buffers": [_p, _p, _p, _p],
    "b2_tendencies": [_p, _p, _p, _p],
    "b2_time_step": [_p, _i, _d, _p, _p],
    "b2_time_step_cfl": [_p, _i, _d, _d, _p, _p, _p, _p],
    "b2_plan_create_slab": [C.POINTER(_p), _i, _i, _i, _d, _d, _d, _i, _i, _i],
    "b2_slab_kept_rows": [_p, C.POINTER(_i)],
    "b2_slab_set_buffers": [_p, _p, _p],
    "b2_slab_set_buffer_strides": [_p, _ll, _ll],
    "b2_slab_buffer_need": [_p, C.POINTER(_ll), C.POINTER(_ll)],
    "b2_set_aliasing": [_p, _i],
    "b2_set_forcing_sparse": [_p, _ll, _p, _p, _i],
    "b2_slab_set_pruning": [_p, _i, _i, _i, _i, _i, _i, _i, _i],
    "b2_slab_curl": [_p, _p, _p],
    "b2_slab_zinv": [_p, _p, _i, _i, _p],
    "b2_slab_set_chunks": [_p, _i],
    "b2_slab_yinv": [_p, _i, _i, _i, _p],
    "b2_slab_xpass": [_p, _i, _p],
    "b2_slab_yfwd": [_p, _i, _i, _i, _p],
    "b2_slab_zfwd": [_p, _i, _i, _p],
    "b2_slab_rk": [_p, _i, _i, _d, _p, _p, _p, _p],
    "b2_slab_phase_a": [_p, _p, _i, _p],
    "b2_slab_phase_b": [_p, _p],
    "b2_slab_phase_c": [_p, _i, _i, _d, _p, _p, _p, _p],
    "b2_nccl_unique_id": [_p],
    "b2_slab_comm_init": [_p, _p],
    "b2_slab_comm_destroy": [_p],
    "b2_slab_tendencies": [_p, _p, _p, _p],
    "b2_slab_time_step": [_p, _i, _d, _p, _p],
    "b2_slab_time_step_cfl": [_p, _i, _d, _d, _p, _p, _p, _p],
    "b2_profile_enable": [_i],
    "b2_profile_reset": [],
    "b2_profile_get": [C.POINTER(_d), C.POINTER(_ll), _i],
    "b2_dev_strided_pass": [_p, _i, _i, _p, _p, _i, _p],
}
_RESTYPES = {"b2_last_error": C.c_char_p, "b2_launch_count": _ll, "b2_observables_size": _ll}

for _name, _args in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError if the symbol is not exported
    _fn.argtypes = _args
    _fn.restype = _RESTYPES.get(_name, _i)

# ns3d.bouss is the stratified kernel without the -N^2 vz coupling (bouss/solver.py:166): N = 0
SOLVER_IDS = {"ns3d": 0, "ns3d.strat": 1, "ns3d.bouss": 1, "ns2d": 2}
SCHEME_IDS = {"RK2": 2, "RK4": 4}


def check(err):
    if err != 0:
        raise B200SpectralError(lib.b2_last_error().decode())


def call(name, *args):
    check(getattr(lib, name)(*args))


def ptr(t):
    """Device pointer of a torch tensor (must be contiguous) or None."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise ValueError("b200spectral needs C-contiguous tensors")
    return t.data_ptr()


def stream_ptr():
    import torch

    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(lib.b2_launch_count())
