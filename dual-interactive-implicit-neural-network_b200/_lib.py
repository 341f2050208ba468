"""ctypes binding of libdiinn_b200.so (C ABI in include/diinn_b200.h). Plain pointers and sizes only.

The library is the product path: if it is missing or cannot be loaded this module raises -- there is no Python or
CPU fallback for any compute entry."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DIINN_B200_LIB") or os.path.join(HERE, "libdiinn_b200.so")  # override: ablation builds

OK = 0
COMPUTE_FP32, COMPUTE_BF16, COMPUTE_FP16, COMPUTE_FP32_SIMT = 0, 1, 2, 3
IO_F32, IO_BF16, IO_BF16_NHWC = 0, 1, 2

STATUS_NAMES = {
    0: "DIINN_OK", -1: "DIINN_ERR_BAD_ARG", -2: "DIINN_ERR_BAD_SHAPE", -3: "DIINN_ERR_BAD_DTYPE",
    -4: "DIINN_ERR_UNSUPPORTED_MODE", -5: "DIINN_ERR_WORKSPACE_TOO_SMALL", -6: "DIINN_ERR_CUDA",
    -7: "DIINN_ERR_NO_WEIGHTS", -8: "DIINN_ERR_UNSUPPORTED_DEVICE",
}

# every symbol include/diinn_b200.h (product surface) and include/diinn_b200_debug.h (taps, probes) declare;
# tests/test_abi.py checks that the .so exports them all and that the lists match the headers
SYMBOLS = [
    "diinn_create", "diinn_destroy", "diinn_last_error", "diinn_set_weights", "diinn_workspace_bytes",
    "diinn_decode", "diinn_decode_multi", "diinn_decode_host", "diinn_query_workspace_bytes", "diinn_query",
    "diinn_query_ensemble", "diinn_set_profiling", "diinn_get_kernel_times", "diinn_launch_count",
    "diinn_version", "diinn_set_output_transform", "diinn_psnr", "diinn_set_bsize", "diinn_set_weights_liif",
]
DEBUG_SYMBOLS = [
    "diinn_debug_gather", "diinn_debug_query_gather", "diinn_debug_set_tap", "diinn_debug_stage_a", "diinn_debug_umma_gemm",
    "diinn_debug_umma_pace", "diinn_debug_read_trace", "diinn_debug_plan_stage_b",
]


class Config(C.Structure):
    _fields_ = [("in_channels", C.c_int), ("hidden", C.c_int), ("n_layers", C.c_int), ("mode", C.c_int),
                ("init_q", C.c_int), ("device", C.c_int)]


class OutputTransform(C.Structure):
    _fields_ = [("affine", C.c_int), ("scale", C.c_float), ("bias", C.c_float), ("clamp", C.c_int), ("lo", C.c_float),
                ("hi", C.c_float), ("quantize_u8", C.c_int)]


class WeightsF32(C.Structure):
    _fields_ = [("k_weight", C.c_void_p * 4), ("k_bias", C.c_void_p * 4), ("q_weight", C.c_void_p * 4),
                ("q_bias", C.c_void_p * 4), ("last_weight", C.c_void_p), ("last_bias", C.c_void_p),
                ("on_device", C.c_int), ("first_weight", C.c_void_p), ("first_bias", C.c_void_p)]


class LiifWeightsF32(C.Structure):
    _fields_ = [("weight", C.c_void_p * 5), ("bias", C.c_void_p * 5), ("on_device", C.c_int)]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). diinn_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i, i64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
    lib.diinn_create.argtypes = [C.POINTER(vp), C.POINTER(Config)]
    lib.diinn_create.restype = i
    lib.diinn_destroy.argtypes = [vp]
    lib.diinn_destroy.restype = None
    lib.diinn_last_error.argtypes = [vp]
    lib.diinn_last_error.restype = C.c_char_p
    lib.diinn_set_weights.argtypes = [vp, C.POINTER(WeightsF32), vp]
    lib.diinn_set_weights.restype = i
    lib.diinn_set_weights_liif.argtypes = [vp, C.POINTER(LiifWeightsF32), vp]
    lib.diinn_set_weights_liif.restype = i
    lib.diinn_workspace_bytes.argtypes = [vp, i, i, i, i, i, i, i, i]
    lib.diinn_workspace_bytes.restype = sz
    lib.diinn_decode.argtypes = [vp, vp, i, i, i, i, i, i, i, i, vp, i64, i64, i64, vp, sz, i, i, vp]
    lib.diinn_decode.restype = i
    lib.diinn_decode_multi.argtypes = [vp, vp, i, i, i, i, i, i, i, i, C.POINTER(vp), i, vp, i64, i64, i64, vp, sz, i, i, vp]
    lib.diinn_decode_multi.restype = i
    lib.diinn_decode_host.argtypes = [vp, vp, i, i, i, i, i, i, i, i, vp, i, i, vp]
    lib.diinn_decode_host.restype = i
    lib.diinn_query_workspace_bytes.argtypes = [vp, i, i, i, i, i]
    lib.diinn_query_workspace_bytes.restype = sz
    lib.diinn_query.argtypes = [vp, vp, i, i, i, i, vp, vp, i, vp, vp, sz, i, i, vp]
    lib.diinn_query.restype = i
    lib.diinn_query_ensemble.argtypes = [vp, vp, i, i, i, i, vp, vp, i, vp, vp, sz, i, i, vp]
    lib.diinn_query_ensemble.restype = i
    lib.diinn_debug_gather.argtypes = [vp, i, i, i, i, vp, vp, vp, vp, vp]
    lib.diinn_debug_gather.restype = i
    lib.diinn_debug_query_gather.argtypes = [vp, i, i, i, vp, vp, i, vp, vp, vp, vp]
    lib.diinn_debug_query_gather.restype = i
    lib.diinn_debug_plan_stage_b.argtypes = [i] * 10 + [vp]
    lib.diinn_debug_plan_stage_b.restype = i
    lib.diinn_debug_set_tap.argtypes = [vp, vp]
    lib.diinn_debug_set_tap.restype = i
    lib.diinn_debug_stage_a.argtypes = [vp, vp, i, i, i, i, vp, vp, sz, i, i, vp]
    lib.diinn_debug_stage_a.restype = i
    lib.diinn_debug_umma_gemm.argtypes = [vp, vp, vp, vp, i, i, i, i, vp]
    lib.diinn_debug_umma_gemm.restype = i
    lib.diinn_debug_read_trace.argtypes = [vp, vp, i]
    lib.diinn_debug_read_trace.restype = i
    lib.diinn_debug_umma_pace.argtypes = [vp, i, i, i, i, vp, i, vp]
    lib.diinn_debug_umma_pace.restype = i
    lib.diinn_set_profiling.argtypes = [vp, i]
    lib.diinn_set_profiling.restype = i
    lib.diinn_get_kernel_times.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                           C.POINTER(i64)]
    lib.diinn_get_kernel_times.restype = i
    lib.diinn_launch_count.argtypes = [vp]
    lib.diinn_launch_count.restype = i64
    lib.diinn_set_output_transform.argtypes = [vp, C.POINTER(OutputTransform)]
    lib.diinn_set_output_transform.restype = i
    lib.diinn_set_bsize.argtypes = [vp, i64]
    lib.diinn_set_bsize.restype = i
    lib.diinn_psnr.argtypes = [vp, vp, vp, i, i, i, i, i, i, i, C.c_float, C.POINTER(C.c_double), vp]
    lib.diinn_psnr.restype = i
    lib.diinn_version.argtypes = []
    lib.diinn_version.restype = C.c_char_p
    _lib = lib
    return lib


class DiinnError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{STATUS_NAMES.get(code, code)}: {msg}")
        self.code = code


def check(lib, handle, code: int):
    if code != OK:
        msg = lib.diinn_last_error(handle)
        raise DiinnError(code, msg.decode() if msg else "")
