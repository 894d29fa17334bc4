"""ctypes binding of libzernike_b200.so (the C ABI in include/zernike_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (``make -C csrc``).  There is
no CPU fallback: if the shared object is missing or no CUDA device is present, calls fail
loudly with ``RuntimeError`` -- nothing silently routes through numpy.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libzernike_b200.so")

# mirror of the #defines in zernike_b200.h
PREC_FP32, PREC_TF32, PREC_TF32X3, PREC_F16, PREC_F16X3 = 0, 1, 2, 3, 4
OUT_REAL, OUT_COMPLEX, OUT_ABS, OUT_ABS_PHASE = 0, 1, 2, 3
NORM_NONE, NORM_L1, NORM_L2, NORM_INF = 0, 1, 2, 3
F32, F64 = 0, 1
ENODEV = -3
ABI_VERSION = 1

PRECISIONS = {"fp32": PREC_FP32, "tf32": PREC_TF32, "tf32x3": PREC_TF32X3, "f16": PREC_F16, "f16x3": PREC_F16X3}

_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p
_i64 = C.c_int64
_int = C.c_int

# name -> (restype, argtypes); every symbol include/zernike_b200.h declares
SIGNATURES = {
    "zb200_abi_version": (_int, []),
    "zb200_last_error": (C.c_char_p, []),
    "zb200_device_info": (_int, [C.POINTER(_int)] * 3),
    "zb200_launch_count": (_i64, []),
    "zb200_reset_launch_count": (None, []),
    "zb200_plan_supports": (_int, [_vp, _int, _int]),
    "zb200_plan_supports_map": (_int, [_vp, _int]),
    "zb200_trim_scratch": (_int, []),
    "zb200_plan_supports_autorange": (_int, [_vp]),
    "zb200_plan_supports_folded_gather": (_int, [_vp]),
    "zb200_num_modes": (_int, [_int]),
    "zb200_num_complex_modes": (_int, [_int]),
    "zb200_mode_table": (_int, [_int, _i32p, _i32p]),
    "zb200_plan_create": (_int, [_int, _int, C.POINTER(_vp)]),
    "zb200_plan_destroy": (None, [_vp]),
    "zb200_plan_n_max": (_int, [_vp]),
    "zb200_plan_size": (_int, [_vp]),
    "zb200_plan_basis_device": (_vp, [_vp]),
    "zb200_plan_basis_to_host": (_int, [_vp, _vp]),
    "zb200_gather_patches_f32": (_int, [_vp, _int, _int, _vp, _i64, _int, _vp, _vp]),
    "zb200_project_patches_f32": (_int, [_vp, _vp, _i64, _int, _int, _vp, _vp, _vp]),
    "zb200_project_patches_ranged_f32": (_int, [_vp, _vp, _i64, C.c_double, _int, _vp, _vp, _vp]),
    "zb200_project_patches_scores_f32": (_int, [_vp, _vp, _i64, _int, _vp, _vp, _int, _int, _vp, _vp]),
    "zb200_project_peaks_f32": (_int, [_vp, _vp, _int, _int, _vp, _i64, _int, _int, _vp, _vp, _vp]),
    "zb200_project_patches_host": (_int, [_vp, _vp, _i64, _int, _vp]),
    "zb200_project_peaks_host": (_int, [_vp, _vp, _int, _int, _int, _vp, _vp, _int, _int, _int, _vp]),
    "zb200_peer_buffer_alloc": (_int, [C.c_size_t, C.POINTER(_vp), _vp]),
    "zb200_peer_buffer_free": (_int, [_vp]),
    "zb200_peer_buffer_open": (_int, [_vp, C.POINTER(_vp)]),
    "zb200_peer_buffer_close": (_int, [_vp]),
    "zb200_project_patches_push_f32": (_int, [_vp, _vp, _i64, _int, _int, C.c_double, _vp, _vp, _int, _vp]),
    "zb200_peer_copy_2d": (_int, [_vp, C.c_size_t, _vp, C.c_size_t, C.c_size_t, C.c_size_t, _vp]),
    "zb200_moment_map_f32": (_int, [_vp, _vp, _int, _int, _int, _int, _int, _vp, _vp]),
    "zb200_symmetry_map_f32": (_int, [_vp, _vp, _int, _int, _int, _int, _int, _vp, _vp, _int, _int, _vp, _vp]),
    "zb200_symmetry_map_host": (_int, [_vp, _vp, _int, _int, _int, _vp, _vp, _int, _int, _vp]),
    "zb200_to_complex": (_int, [_int, _vp, _i64, _i64, _i64, _vp, _vp, _int, _vp, _i64, _i64, _vp]),
    "zb200_to_real": (_int, [_int, _vp, _i64, _i64, _i64, _vp, _vp, _int, _vp, _i64, _i64, _vp]),
    "zb200_select_modes": (_int, [_int, _int, _vp, _i64, _i64, _i64, _vp, _int, _vp, _i64, _i64, _vp]),
    "zb200_normalize": (_int, [_int, _int, _vp, _i64, _int, _i64, _i64, _int, C.c_double, _vp, _vp]),
    "zb200_rotate": (_int, [_int, _vp, _i64, _int, _i64, _i64, _vp, C.c_double, _vp, _vp]),
    "zb200_rot_scores": (_int, [_int, _vp, _i64, _int, _i64, _i64, _vp, _vp, _int, _int, _vp, _i64, _i64, _vp]),
    "zb200_mirror_scores": (_int, [_int, _vp, _i64, _int, _i64, _i64, _vp, _vp, _int, _vp, _vp]),
    "zb200_complex_abs_phase": (_int, [_int, _vp, _i64, _vp, _vp, _vp]),
    "zb200_render_atoms_f32": (_int, [_vp, _vp, C.c_double, _i64, C.c_double, C.c_double, _int, _int, _vp, _int, _vp]),
    "zb200_lattice_coords_f64": (_int, [_int, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_double, _vp, _vp, _vp, _vp, _vp]),
    "zb200_render_stamps_f32": (_int, [_vp, _vp, _i64, C.c_double, C.c_double, _int, _int, _int, _vp, _int, _vp]),
    "zb200_local_max_f32": (_int, [_vp, _int, _int, C.c_double, _int, C.c_double, _vp, _i64, C.POINTER(_i64),
                                   C.POINTER(_i64), _vp]),
    "zb200_gram_f32": (_int, [_vp, _i64, _int, _vp, _vp, _vp]),
    "zb200_pca_scores_f32": (_int, [_vp, _i64, _int, _vp, _vp, _int, _vp, _vp]),
    "zb200_kmeans_mindist_f32": (_int, [_vp, _i64, _int, _vp, _vp, _int, _vp, _vp, _vp, _vp]),
    "zb200_kmeans_step_f32": (_int, [_vp, _i64, _int, _vp, _vp, _int, _vp, _int, _vp, _vp, C.POINTER(_i64), _vp]),
    "zb200_gmm_estep_f32": (_int, [_vp, _i64, _int, _vp, _vp, _vp, _vp, _int, _vp, _vp, C.POINTER(C.c_double), _vp]),
    "zb200_gmm_mstep_f32": (_int, [_vp, _i64, _int, _vp, _vp, _vp, _int, _vp, _vp, _vp, _vp]),
    "zb200_overlap_add_f64": (_int, [_vp, _vp, _int, _vp, _int, _int, _int, _int, _int, _vp, _vp]),
    "zb200_download_as_f64": (_int, [_vp, _i64, _vp, _vp]),
    "zb200_cast": (_int, [_int, _vp, _int, _vp, _i64, _vp]),
}

_lock = threading.Lock()
_lib = None


class ZernikeB200Error(RuntimeError):
    """A C-ABI call failed; the message is zb200_last_error()."""


def load():
    """Load (once) and return the ctypes handle.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C motif-learn_b200/csrc`). motif_learn_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        got = lib.zb200_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError(f"libzernike_b200.so ABI {got} != expected {ABI_VERSION}; rebuild")
        _lib = lib
        return lib


def last_error() -> str:
    return load().zb200_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc < 0:
        raise ZernikeB200Error(f"{what + ': ' if what else ''}{last_error()} (code {rc})")


def require_cuda():
    """torch is the device-memory/stream plumbing; fail loudly when no GPU is present."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("motif_learn_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def current_stream_ptr() -> int:
    import torch
    return int(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    return int(load().zb200_launch_count())


def reset_launch_count() -> None:
    load().zb200_reset_launch_count()
