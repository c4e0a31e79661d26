"""ctypes binding of libdroneyolo.so (include/droneyolo.h).

There is no CPU fallback: every op raises if the library is missing or the tensors are not on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "libdroneyolo.so"
if os.environ.get("DY_LIB"):                      # developer switch: a debug build (tools/trace_conv.py)
    LIB_PATH = Path(os.environ["DY_LIB"])

DY_BF16, DY_F32, DY_U8 = 0, 1, 2
DY_ACT_NONE, DY_ACT_SILU = 0, 1
DY_NHWC, DY_NCHW = 0, 1


class DroneYoloError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("in_", C.c_void_p), ("in_ld", C.c_int32),
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32),
        ("weight", C.c_void_p), ("bias", C.c_void_p),
        ("Cout", C.c_int32), ("ksize", C.c_int32), ("stride", C.c_int32),
        ("out", C.c_void_p), ("out_ld", C.c_int32), ("out_dtype", C.c_int32),
        ("residual", C.c_void_p), ("res_ld", C.c_int32),
        ("act", C.c_int32),
        ("up_out", C.c_void_p), ("up_ld", C.c_int32),
        ("weight2", C.c_void_p), ("bias2", C.c_void_p),
        ("Cout2", C.c_int32), ("out2", C.c_void_p), ("out2_ld", C.c_int32),
        ("tail_decode", C.c_int32), ("y", C.c_void_p), ("y_A", C.c_int32), ("y_nc", C.c_int32), ("y_anchor_off", C.c_int32),
        ("y_stride", C.c_float),
        ("pre_add", C.c_void_p), ("pre_ld", C.c_int32),
    ]


class DecodeDesc(C.Structure):
    _fields_ = [
        ("lvl", C.c_void_p * 4), ("ld", C.c_int32 * 4), ("H", C.c_int32 * 4), ("W", C.c_int32 * 4),
        ("stride", C.c_float * 4),
        ("nl", C.c_int32), ("B", C.c_int32), ("nc", C.c_int32), ("dtype", C.c_int32), ("layout", C.c_int32),
        ("out", C.c_void_p),
        ("A_total", C.c_int32), ("anchor_off", C.c_int32 * 4),
    ]


class NmsDesc(C.Structure):
    _fields_ = [
        ("pred", C.c_void_p), ("B", C.c_int32), ("nc", C.c_int32), ("A", C.c_int32),
        ("conf_thres", C.c_float),
        ("iou_thres", C.c_double),
        ("max_det", C.c_int32), ("max_nms", C.c_int32), ("max_wh", C.c_float),
        ("agnostic", C.c_int32), ("multi_label", C.c_int32),
        ("classes_host", C.POINTER(C.c_int32)), ("n_classes", C.c_int32),
        ("xyxy_in_place", C.c_int32),
        ("out", C.c_void_p), ("counts", C.c_void_p), ("kept", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("rescale", C.c_void_p),
    ]


# name -> (restype, argtypes): every symbol include/droneyolo.h declares
SYMBOLS = {
    "dy_version": (C.c_int, []),
    "dy_last_error": (C.c_char_p, []),
    "dy_device_check": (C.c_int, [C.c_int]),
    "dy_conv2d": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    "dy_stem_conv": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "dy_sppf_pool": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "dy_letterbox_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "dy_letterbox_u8_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "dy_upsample2x": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "dy_dwconv3x3s2": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "dy_detect_decode": (C.c_int, [C.POINTER(DecodeDesc), C.c_void_p]),
    "dy_nms_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "dy_nms": (C.c_int, [C.POINTER(NmsDesc), C.c_void_p]),
    "dy_box_nms_f64_workspace_bytes": (C.c_size_t, [C.c_int]),
    "dy_box_nms_f64": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "dy_program_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "dy_program_destroy": (None, [C.c_void_p]),
    "dy_program_add_conv": (C.c_int, [C.c_void_p, C.POINTER(ConvDesc)]),
    "dy_program_add_stem": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "dy_program_add_sppf_pool": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "dy_program_add_upsample2x": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "dy_program_add_dwconv3x3s2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "dy_program_add_decode": (C.c_int, [C.c_void_p, C.POINTER(DecodeDesc)]),
    "dy_program_add_nms": (C.c_int, [C.c_void_p, C.POINTER(NmsDesc)]),
    "dy_program_set_lane": (C.c_int, [C.c_void_p, C.c_int]),
    "dy_program_add_sync": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "dy_program_run": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    "dy_program_num_launches": (C.c_int, [C.c_void_p]),
    "dy_program_profile": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_int]),
    "dy_program_num_ops": (C.c_int, [C.c_void_p]),
    "dy_selftest_umma": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_float), C.c_void_p]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the library once; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise DroneYoloError(
                f"{LIB_PATH} is missing: build it with `python -m drone_yolo_b200.build` (nvcc, sm_100a). "
                "drone_yolo_b200 has no CPU or PyTorch fallback for its kernels."
            )
        handle = C.CDLL(str(LIB_PATH))
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().dy_last_error().decode(errors="replace")
        raise DroneYoloError(f"{what or 'libdroneyolo'} failed ({rc}): {msg}")


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise DroneYoloError(
                "drone_yolo_b200 kernels run on CUDA (sm_100a) tensors only; got a tensor on "
                f"{t.device}. There is no CPU fallback."
            )
