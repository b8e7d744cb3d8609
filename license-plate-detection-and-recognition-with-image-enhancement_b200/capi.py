"""ctypes binding of the C ABI in include/lpsr_b200.h (liblpsr_b200.so, built in-tree by build_ext.py).

This is the only place the Python host touches native code.  There is no fallback: if the shared library is
missing or no sm_100 GPU is present every call raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# LPSR_B200_LIB: another build of the same library (A/B measurements of kernel variants); default: the in-tree build
LIB_PATH = os.environ.get("LPSR_B200_LIB") or os.path.join(HERE, "liblpsr_b200.so")
ABI_VERSION = 1

PREC = {"fp32": 0, "bf16": 1, "fp16": 2}

EXPORTS = [
    "lpsr_create", "lpsr_destroy", "lpsr_load_weights", "lpsr_num_live_tensors", "lpsr_live_tensor_name",
    "lpsr_live_tensor_numel", "lpsr_output_shape", "lpsr_workspace_bytes", "lpsr_forward", "lpsr_forward_profiled", "lpsr_forward_host",
    "lpsr_forward_launch_count", "lpsr_debug_read_tap", "lpsr_debug_umma_trace", "lpsr_last_error", "lpsr_abi_version", "lpsr_device_sm",
    "lpsr_op_pixel_unshuffle2", "lpsr_op_pixel_shuffle2", "lpsr_op_conv2d", "lpsr_preprocess_resize",
    "lpsr_op_yolo_nms_workspace_bytes", "lpsr_op_yolo_nms",
]


class LpsrConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("num_channels", C.c_int32),
                ("num_features", C.c_int32), ("growth_rate", C.c_int32), ("num_blocks", C.c_int32),
                ("num_layers", C.c_int32), ("out_channels", C.c_int32), ("precision", C.c_int32),
                ("reserved", C.c_int32 * 7)]


class LpsrTensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("numel", C.c_int64), ("on_device", C.c_int32),
                ("reserved", C.c_int32)]


class LpsrError(RuntimeError):
    pass


_lib = None


def load_library() -> C.CDLL:
    """Load liblpsr_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LpsrError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU / PyTorch fallback for the LPSR forward)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.lpsr_create.argtypes = [C.POINTER(vp), C.POINTER(LpsrConfig)]
    lib.lpsr_destroy.argtypes = [vp]
    lib.lpsr_load_weights.argtypes = [vp, C.POINTER(LpsrTensorDesc), i32]
    lib.lpsr_num_live_tensors.argtypes = [vp]
    lib.lpsr_live_tensor_name.argtypes = [vp, i32]
    lib.lpsr_live_tensor_name.restype = C.c_char_p
    lib.lpsr_live_tensor_numel.argtypes = [vp, i32]
    lib.lpsr_live_tensor_numel.restype = i64
    lib.lpsr_output_shape.argtypes = [vp, i32, i32, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    lib.lpsr_workspace_bytes.argtypes = [vp, i32, i32, i32]
    lib.lpsr_workspace_bytes.restype = C.c_size_t
    lib.lpsr_forward.argtypes = [vp, vp, vp, i32, i32, i32, vp, C.c_size_t, vp]
    lib.lpsr_forward_profiled.argtypes = [vp, vp, vp, i32, i32, i32, vp, C.c_size_t, vp, vp, vp, i32, i32]
    lib.lpsr_forward_host.argtypes = [vp, vp, vp, i32, i32, i32]
    lib.lpsr_forward_launch_count.argtypes = [vp, i32, i32, i32]
    lib.lpsr_debug_read_tap.argtypes = [vp, C.c_char_p, vp, i64, i32, i32, i32, vp, vp]
    lib.lpsr_last_error.argtypes = [vp]
    lib.lpsr_last_error.restype = C.c_char_p
    lib.lpsr_abi_version.argtypes = []
    lib.lpsr_device_sm.argtypes = [vp]
    lib.lpsr_preprocess_resize.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp]
    lib.lpsr_op_pixel_unshuffle2.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.lpsr_op_pixel_shuffle2.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.lpsr_op_conv2d.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.lpsr_op_yolo_nms_workspace_bytes.argtypes = [i32, i32, i32, i32]
    lib.lpsr_op_yolo_nms_workspace_bytes.restype = C.c_size_t
    lib.lpsr_op_yolo_nms.argtypes = [vp, vp, i32, i32, i32, C.c_float, C.c_float, vp, i32, i32, i32, i32, i32, vp, vp, vp, C.c_size_t, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int:   # default
            fn.restype = C.c_int
    if lib.lpsr_abi_version() != ABI_VERSION:
        raise LpsrError(f"ABI mismatch: library {lib.lpsr_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, handle=None, what: str = "lpsr") -> None:
    """Integer status -> RuntimeError, the way torch ops fail at the reference call sites."""
    if rc == 0:
        return
    lib = load_library()
    msg = lib.lpsr_last_error(handle)
    raise LpsrError(f"{what} failed (status {rc}): {msg.decode() if msg else '?'}")
