"""ctypes binding of libhifidiff_b200.so (the C ABI in include/hifidiff_b200.h).

There is no fallback: if the shared library is missing the import of any compute entry point
raises, and if it is present but no sm_100 GPU is visible `hd_create` fails with
HD_ERR_UNSUPPORTED.  The library is built in-tree by `__graft_entry__.build()`
(`make -C hifidiff_b200/csrc`).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libhifidiff_b200.so")

HD_OK = 0
HD_MODEL_DENOISER = 0
HD_MODEL_FUSED = 1
HD_PRECISION_BF16 = 0
HD_PRECISION_FP32 = 1

# every symbol include/hifidiff_b200.h declares
EXPORTED_SYMBOLS = (
    "hd_abi_version", "hd_create", "hd_destroy", "hd_last_error", "hd_get_info", "hd_load_weights",
    "hd_set_time_frequencies", "hd_set_condition", "hd_denoise_step", "hd_denoise_step_taps",
    "hd_sample", "hd_sampler_update", "hd_debug_gemm", "hd_synchronize", "hd_profile_step", "hd_debug_gemm_trace", "hd_load_fpg_weights", "hd_fpg_forward", "hd_load_idc_weights", "hd_idc_forward", "hd_load_cr_weights", "hd_cr_forward", "hd_debug_gemm_time",
)


class HdConfig(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("model", C.c_int32), ("precision", C.c_int32),
                ("latent_size", C.c_int32), ("device", C.c_int32), ("max_batch", C.c_int32),
                ("max_steps", C.c_int32), ("use_graph", C.c_int32)]


class HdTensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("dtype", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int64 * 4)]


class HdStepCoef(C.Structure):
    _fields_ = [("timestep", C.c_float), ("sqrt_beta_prod", C.c_float), ("sqrt_alpha_prod", C.c_float),
                ("clip", C.c_float), ("k_x0", C.c_float), ("k_eps", C.c_float), ("k_x", C.c_float),
                ("k_noise", C.c_float)]


class HdInfo(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("abi_version", C.c_int32), ("sm_major", C.c_int32),
                ("sm_minor", C.c_int32), ("sm_count", C.c_int32), ("launches_per_step", C.c_int32),
                ("weight_bytes", C.c_int64), ("workspace_bytes", C.c_int64),
                ("weight_elems_per_step", C.c_int64), ("flops_per_face_step", C.c_double)]


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Loads the shared library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(hifidiff_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    lib.hd_abi_version.restype = i32
    lib.hd_abi_version.argtypes = []
    lib.hd_create.restype = i32
    lib.hd_create.argtypes = [C.POINTER(vp), C.POINTER(HdConfig)]
    lib.hd_destroy.restype = None
    lib.hd_destroy.argtypes = [vp]
    lib.hd_last_error.restype = C.c_char_p
    lib.hd_last_error.argtypes = [vp]
    lib.hd_get_info.restype = i32
    lib.hd_get_info.argtypes = [vp, C.POINTER(HdInfo)]
    lib.hd_load_weights.restype = i32
    lib.hd_load_weights.argtypes = [vp, C.POINTER(HdTensorDesc), i32, vp]
    lib.hd_set_time_frequencies.restype = i32
    lib.hd_set_time_frequencies.argtypes = [vp, vp]
    lib.hd_set_condition.restype = i32
    lib.hd_set_condition.argtypes = [vp, C.POINTER(vp), vp, i32, vp]
    lib.hd_denoise_step.restype = i32
    lib.hd_denoise_step.argtypes = [vp, vp, vp, i32, vp, i32, vp]
    lib.hd_denoise_step_taps.restype = i32
    lib.hd_denoise_step_taps.argtypes = [vp, vp, vp, i32, vp, i32, C.POINTER(C.c_char_p), C.POINTER(vp), i32, vp]
    lib.hd_sample.restype = i32
    lib.hd_sample.argtypes = [vp, vp, C.POINTER(HdStepCoef), i32, u64, i64, i32, vp, vp]
    lib.hd_sampler_update.restype = i32
    lib.hd_sampler_update.argtypes = [vp, vp, vp, C.POINTER(HdStepCoef), i32, u64, i64, i32, vp, vp]
    lib.hd_debug_gemm.restype = i32
    lib.hd_debug_gemm.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.hd_profile_step.restype = i32
    lib.hd_profile_step.argtypes = [vp, i32, i32, vp, vp, i32, i32, C.POINTER(i32)]
    lib.hd_debug_gemm_trace.restype = i32
    lib.hd_debug_gemm_trace.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, i32, C.POINTER(i32), C.POINTER(i32)]
    lib.hd_load_fpg_weights.restype = i32
    lib.hd_load_fpg_weights.argtypes = [vp, C.POINTER(HdTensorDesc), i32, vp]
    lib.hd_fpg_forward.restype = i32
    lib.hd_fpg_forward.argtypes = [vp, vp, C.POINTER(vp), i32, vp]
    lib.hd_load_idc_weights.restype = i32
    lib.hd_load_idc_weights.argtypes = [vp, C.POINTER(HdTensorDesc), i32, vp]
    lib.hd_idc_forward.restype = i32
    lib.hd_idc_forward.argtypes = [vp, vp, i32, vp, i32, vp]
    lib.hd_load_cr_weights.restype = i32
    lib.hd_load_cr_weights.argtypes = [vp, C.POINTER(HdTensorDesc), i32, vp]
    lib.hd_cr_forward.restype = i32
    lib.hd_cr_forward.argtypes = [vp, vp, i32, vp, i32, vp]
    lib.hd_debug_gemm_time.restype = i32
    lib.hd_debug_gemm_time.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, C.POINTER(C.c_float)]
    lib.hd_synchronize.restype = i32
    lib.hd_synchronize.argtypes = [vp]
    _lib = lib
    return lib


def check(handle, status: int, what: str) -> None:
    if status != HD_OK:
        msg = load().hd_last_error(handle)
        raise RuntimeError(f"{what} failed (hd_status {status}): {msg.decode() if msg else '?'}")
