"""ctypes binding of libsdnq_b200.so (the C ABI declared in include/sdnq_b200.h).

There is deliberately no fallback: if the shared object is missing or the device is not sm_100 every kernel
entry point raises.  Build it with `python sdnq_b200/csrc/build.py` (or `__graft_entry__.build()`).
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsdnq_b200.so")

SDNQ_F32, SDNQ_BF16, SDNQ_F16, SDNQ_I8, SDNQ_U8, SDNQ_F8E4M3, SDNQ_I32, SDNQ_F8E5M2 = range(8)
SDNQ_W_INT, SDNQ_W_MINIFLOAT, SDNQ_W_FP8_E4M3FN, SDNQ_W_FP8_E5M2 = range(4)
ABI_VERSION = 2


class Conv2dGeometry(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in ("batch", "channels", "height", "width", "x_stride_b", "x_stride_c", "x_stride_h", "x_stride_w")] + \
               [(n, ctypes.c_int32) for n in ("kernel_h", "kernel_w", "stride_h", "stride_w", "pad_h", "pad_w", "dilation_h", "dilation_w")]


class WeightFormat(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("bits", ctypes.c_int32), ("is_unsigned", ctypes.c_int32),
                ("exponent", ctypes.c_int32), ("mantissa", ctypes.c_int32), ("word_bytes", ctypes.c_int32)]


class DequantJob(ctypes.Structure):
    _fields_ = [("weight", ctypes.c_void_p), ("fmt", WeightFormat), ("scale", ctypes.c_void_p), ("zero_point", ctypes.c_void_p),
                ("N", ctypes.c_int64), ("K", ctypes.c_int64), ("group_size", ctypes.c_int64),
                ("svd_up", ctypes.c_void_p), ("up_stride_n", ctypes.c_int64), ("up_stride_r", ctypes.c_int64),
                ("svd_down", ctypes.c_void_p), ("down_stride_r", ctypes.c_int64), ("down_stride_k", ctypes.c_int64),
                ("svd_rank", ctypes.c_int32), ("svd_dtype", ctypes.c_int32), ("out", ctypes.c_void_p),
                ("out_dtype", ctypes.c_int32), ("reserved", ctypes.c_int32)]


_P, _I, _L, _Z = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t
_WF = ctypes.POINTER(WeightFormat)

# name -> (restype, argtypes); must list every symbol include/sdnq_b200.h declares (tests/test_c_abi.py checks)
SIGNATURES = {
    "sdnq_b200_abi_version": (_I, []),
    "sdnq_b200_last_error": (ctypes.c_char_p, []),
    "sdnq_b200_check_device": (_I, [_I]),
    "sdnq_b200_unpack": (_I, [_P, _WF, _P, _I, _L, _P]),
    "sdnq_b200_dequant": (_I, [_P, _WF, _P, _P, _I, _L, _L, _L, _P, _L, _L, _P, _L, _L, _I, _I, _I, _P, _I, _P]),
    "sdnq_b200_dequant_nd": (_I, [_P, _WF, _P, _P, _I, _I, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64), _P, _I, _P, _I, _P]),
    "sdnq_b200_dequant_batch_table_bytes": (_Z, [_I]),
    "sdnq_b200_dequant_batch_plan": (_I, [_P, _I, _P, _P]),
    "sdnq_b200_dequant_batch_run": (_I, [_P, _P, _P]),
    "sdnq_b200_quantize_weight": (_I, [_P, _I, _L, _L, _L, _WF, _I, _P, _P, _P, _P]),
    "sdnq_b200_embedding": (_I, [_P, _WF, _P, _P, _I, _L, _L, _L, _P, _L, _L, _P, _L, _L, _I, _I, _I, ctypes.POINTER(ctypes.c_int64), _L, ctypes.c_float, _P, _I, _P]),
    "sdnq_b200_requant": (_I, [_P, _WF, _P, _P, _I, _L, _L, _L, _I, _P, _P, _P, _P, _P]),
    "sdnq_b200_act_quant": (_I, [_P, _I, _L, _L, _L, _I, _I, _P, _P, _P, _P, _P, _P]),
    "sdnq_b200_conv_act_quant": (_I, [_P, _I, ctypes.POINTER(Conv2dGeometry), _I, _I, _P, _P, _P, _P, _P, _P]),
    "sdnq_b200_conv_act_quant_workspace_bytes": (_Z, [ctypes.POINTER(Conv2dGeometry), _I]),
    "sdnq_b200_conv_act_quant_ws": (_I, [_P, _I, ctypes.POINTER(Conv2dGeometry), _I, _I, _P, _P, _P, _P, _P, _P, _Z, _P]),
    "sdnq_b200_rows_to_nchw": (_I, [_P, _P, _I, _L, _L, _L, _P]),
    "sdnq_b200_scaled_mm": (_I, [_P, _P, _I, _P, _P, _P, _I, _L, _P, _P, _P, _P, _P, _I, _L, _L, _L, _P]),
    "sdnq_b200_scaled_mm_workspace_bytes": (_Z, []),
    "sdnq_b200_attention_workspace_bytes": (_Z, [_L, _L, _L, _L]),
    "sdnq_b200_attention": (_I, [_P, _P, _P, _I, _I, _P, _P, _P, _P, _I, ctypes.POINTER(ctypes.c_int64), _P, _P, _I, _L, _L, _L, _L, _L, _L, _L, _L,
                                 ctypes.c_float, _I, _P, _Z, _P]),
    "sdnq_b200_attn_colmean": (_I, [_P, _I, _L, _L, _L, _P, _P]),
    "sdnq_b200_attn_quant": (_I, [_P, _I, _L, _L, _P, _L, _I, _P, _P, _P]),
    "sdnq_b200_smooth_k": (_I, [_P, _I, _L, _L, _L, _P, _I, _P]),
    "sdnq_b200_scaled_mm_ws": (_I, [_P, _P, _I, _P, _P, _P, _I, _L, _P, _P, _P, _P, _P, _I, _L, _L, _L, _P, _Z, _P]),
    "sdnq_b200_scaled_mm_packed": (_I, [_P, _P, _WF, _P, _P, _P, _I, _L, _P, _P, _P, _I, _L, _L, _L, _P]),
    "sdnq_b200_mm": (_I, [_P, _P, _I, _P, _L, _L, _L, _P]),
    "sdnq_b200_linear_small_m": (_I, [_P, _I, _L, _P, _I, _P, _P, _P, _I, _P, _L, _L, _L, _P]),
    "sdnq_b200_linear_small_m_packed": (_I, [_P, _I, _L, _P, _WF, _P, _P, _L, _P, _I, _L, _P, _L, _L, _L, _P]),
    "sdnq_b200_linear_w4a16": (_I, [_P, _I, _L, _P, _WF, _P, _P, _L, _P, _P, _I, _P, _I, _P, _L, _L, _L, _P]),
    "sdnq_b200_svd_low": (_I, [_P, _I, _L, _P, _I, _P, _L, _L, _P]),
    "sdnq_b200_scaled_mm_svd": (_I, [_P, _P, _I, _WF, _P, _P, _P, _I, _L, _P, _P, _P, _P, _P, _P, _I, _I, _P, _I, _L, _L, _L, _P]),
    "sdnq_b200_scaled_mm_grouped": (_I, [_P, _P, _I, _WF, _P, _P, _P, _I, _P, _P, _P, _P, _I, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64),
                                         ctypes.POINTER(ctypes.c_void_p), _I, _L, _L, _P]),
    "sdnq_b200_linear_w8a8_workspace_bytes": (_Z, [_L, _L]),
    "sdnq_b200_linear_w8a8": (_I, [_P, _I, _L, _P, _I, _P, _P, _P, _P, _I, _I, _P, _I, _L, _L, _L, _P, _Z, _P]),
    "sdnq_b200_linear_w8a8_fused": (_I, [_P, _I, _L, _P, _I, _P, _P, _I, _P, _I, _L, _L, _L, _P, _Z, _P]),
    "sdnq_b200_stream_capture_id": (_L, [_P]),
    "sdnq_b200_launch_count": (_L, [_I]),
}

_lib = None
_lock = threading.Lock()


class SDNQKernelError(RuntimeError):
    pass


def load():
    """Load the shared object (once).  Raises if it has not been built -- there is no CPU / eager fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise SDNQKernelError(
                f"{LIB_PATH} is missing: the sm_100a kernels have not been built. Run `python sdnq_b200/csrc/build.py`; "
                "sdnq_b200 has no CPU or eager fallback for the quantized-Linear path.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.sdnq_b200_abi_version() != ABI_VERSION:
            raise SDNQKernelError(f"ABI mismatch: library {lib.sdnq_b200_abi_version()} vs binding {ABI_VERSION}; rebuild")
        _lib = lib
    return _lib


def check(rc: int):
    if rc != 0:
        msg = load().sdnq_b200_last_error().decode(errors="replace")
        raise SDNQKernelError(f"sdnq_b200 kernel call failed ({rc}): {msg}")


def launch_count(reset: bool = False) -> int:
    return int(load().sdnq_b200_launch_count(1 if reset else 0))
