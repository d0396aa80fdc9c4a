"""ctypes binding of libposecodec.so (include/posecodec.h).

This module is the thin host-side layer of the graft: it loads the shared
library, mirrors the parameter structs, extracts raw device pointers from
torch / DLPack / ``__cuda_array_interface__`` objects and maps status codes to
the exception types the reference uses (``ValueError`` for configuration
errors, ``RuntimeError`` for CUDA failures).  There is no CPU fallback: if the
library is missing or no sm_100 device is visible the codec raises.
"""
import ctypes
import os
from ctypes import (
    POINTER,
    Structure,
    c_char_p,
    c_double,
    c_float,
    c_int,
    c_int32,
    c_int64,
    c_uint8,
    c_void_p,
)

PC_MAX_JOINTS = 64
PC_NMS_MAX_PEOPLE = 1024
PC_MAX_DARK_KERNEL = 17
PC_MAX_GROUPS = 128
PC_MAX_DETECTIONS = 64
PC_MAX_SCALES = 4

PC_OK = 0
PC_ERR_INVALID_ARGUMENT = -1
PC_ERR_UNSUPPORTED = -2
PC_ERR_CUDA = -3
PC_ERR_NO_DEVICE = -4

_HERE = os.path.dirname(os.path.abspath(__file__))
# POSECODEC_LIB: development override used to A/B two builds of the library in one session
# (scripts/gpu_ab.sh); it must still be a libposecodec build -- every symbol is checked on load.
LIB_PATH = os.environ.get("POSECODEC_LIB") or os.path.join(_HERE, "csrc", "libposecodec.so")


class BoxParams(Structure):
    _fields_ = [
        ("image_w", c_int32),
        ("image_h", c_int32),
        ("pixel_std", c_float),
        ("scale_padding", c_float),
    ]


class AffineParams(Structure):
    _fields_ = [
        ("image_w", c_int32),
        ("image_h", c_int32),
        ("pixel_std", c_float),
        ("use_udp", c_int32),
    ]


class WarpParams(Structure):
    _fields_ = [("dst_w", c_int32), ("dst_h", c_int32), ("channels", c_int32)]


class WarpNormParams(Structure):
    _fields_ = [
        ("dst_w", c_int32),
        ("dst_h", c_int32),
        ("channels", c_int32),
        ("mean", c_float * 4),
        ("std", c_float * 4),
    ]


class EncodeParams(Structure):
    _fields_ = [
        ("num_joints", c_int32),
        ("image_w", c_int32),
        ("image_h", c_int32),
        ("heatmap_w", c_int32),
        ("heatmap_h", c_int32),
        ("sigma", c_float),
        ("use_udp", c_int32),
        ("use_joint_weights", c_int32),
        ("joint_weights", c_float * PC_MAX_JOINTS),
    ]


class TopDownDecodeParams(Structure):
    _fields_ = [
        ("num_joints", c_int32),
        ("height", c_int32),
        ("width", c_int32),
        ("pixel_std", c_float),
        ("to_original", c_int32),
        ("shift_coordinate", c_int32),
        ("use_udp", c_int32),
        ("dark_udp_refine", c_int32),
        ("kernel_size", c_int32),
        ("flip_test", c_int32),
        ("shift_heatmap", c_int32),
        ("flip_index", c_int32 * PC_MAX_JOINTS),
        ("dark_kernel_set", c_int32),
        ("dark_kernel", c_float * (PC_MAX_DARK_KERNEL * PC_MAX_DARK_KERNEL)),
    ]


class BottomUpDecodeParams(Structure):
    _fields_ = [
        ("num_joints", c_int32),
        ("num_stages", c_int32),
        ("h0", c_int32),
        ("w0", c_int32),
        ("h1", c_int32),
        ("w1", c_int32),
        ("mask_h", c_int32),
        ("mask_w", c_int32),
        ("use_nms", c_int32),
        ("nms_kernel", c_int32),
        ("max_num", c_int32),
        ("shift_coordinate", c_int32),
        ("tag_per_joint", c_int32),
    ]


class GroupParams(Structure):
    _fields_ = [
        ("num_joints", c_int32),
        ("max_num", c_int32),
        ("vis_thr", c_float),
        ("tag_thr", c_float),
        ("ignore_too_much", c_int32),
        ("use_rounded_norm", c_int32),
        ("joint_order", c_int32 * PC_MAX_JOINTS),
        ("max_groups", c_int32),
    ]


class BottomUpEncodeParams(Structure):
    _fields_ = [
        ("num_joints", c_int32),
        ("num_scales", c_int32),
        ("num_people", c_int32),
        ("max_num", c_int32),
        ("heatmap_w", c_int32 * PC_MAX_SCALES),
        ("heatmap_h", c_int32 * PC_MAX_SCALES),
        ("sigma", c_float),
        ("tag_per_joint", c_int32),
    ]


class RefineParams(Structure):
    _fields_ = [("num_joints", c_int32), ("height", c_int32), ("width", c_int32),
                ("max_groups", c_int32)]


class OksNmsParams(Structure):
    _fields_ = [
        ("num_joints", c_int32),
        ("rescore", c_int32),
        ("use_nms", c_int32),
        ("soft", c_int32),
        ("max_dets", c_int32),
        ("use_iou_vis_thr", c_int32),
        ("rescore_vis_thr", c_float),
        ("oks_thr", c_float),
        ("iou_vis_thr", c_float),
        ("max_people_per_image", c_int32),
        ("sigmas", ctypes.c_double * PC_MAX_JOINTS),
        ("rescore_vis_thr_f64", c_double),
        ("iou_vis_thr_f64", c_double),
    ]


class GatherTarget(Structure):
    _fields_ = [
        ("h_peer_tables", POINTER(c_void_p)),
        ("num_peers", c_int32),
        ("d_multicast_table", c_void_p),
        ("row_offset", c_int64),
        ("h_peer_flags", POINTER(c_void_p)),
        ("num_flag_peers", c_int32),
        ("my_rank", c_int32),
        ("d_step", c_void_p),
        ("d_counter", c_void_p),
    ]


class AffineHostParams(Structure):
    _fields_ = [
        ("src_h", c_int32),
        ("src_w", c_int32),
        ("channels", c_int32),
        ("image_w", c_int32),
        ("image_h", c_int32),
        ("pixel_std", c_float),
        ("scale_padding", c_float),
        ("use_udp", c_int32),
        ("upload", c_int32),
    ]


UPLOAD_FULL, UPLOAD_ROI, UPLOAD_ROI_KERNEL = 0, 1, 2   # PC_UPLOAD_* of posecodec.h

# name -> (restype, argtypes); one entry per function declared in posecodec.h
_P = c_void_p
SIGNATURES = {
    "pc_version": (c_int, []),
    "pc_last_error": (c_char_p, []),
    "pc_device_info": (c_int, [c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "pc_box_to_center_scale": (c_int, [_P, _P, _P, POINTER(BoxParams), c_int64, _P]),
    "pc_affine_matrices": (c_int, [_P, _P, _P, _P, _P, POINTER(AffineParams), c_int64, _P]),
    "pc_affine_from_points": (c_int, [_P, _P, _P, _P, c_int64, _P]),
    "pc_invert_affine": (c_int, [_P, _P, c_int64, _P]),
    "pc_warp_affine_u8": (c_int, [_P, _P, _P, _P, _P, POINTER(WarpParams), c_int64, _P]),
    "pc_warp_affine_u8_norm_chw": (
        c_int, [_P, _P, _P, _P, _P, POINTER(WarpNormParams), c_int64, _P]),
    "pc_rescale_pad_u8": (
        c_int, [_P, _P, _P, _P, _P, _P, c_int32, c_int32, c_int32, c_int64, _P]),
    "pc_rescale_pad_u8_norm_chw": (
        c_int, [_P, _P, _P, _P, _P, _P, POINTER(WarpNormParams), c_int64, _P]),
    "pc_affine_joints": (c_int, [_P, _P, c_int32, c_int32, c_int64, _P]),
    "pc_topdown_encode": (c_int, [_P, _P, _P, POINTER(EncodeParams), c_int64, _P]),
    "pc_topdown_decode": (
        c_int,
        [_P, _P, _P, _P, _P, _P, _P, POINTER(TopDownDecodeParams), c_int64, _P],
    ),
    "pc_bottomup_encode": (c_int, [_P, _P, _P, POINTER(BottomUpEncodeParams), c_int64, _P]),
    "pc_bottomup_decode": (
        c_int,
        [_P, _P, _P, _P, _P, _P, _P, _P, POINTER(BottomUpDecodeParams), c_int64, _P],
    ),
    "pc_bottomup_decode_stats": (c_int, [POINTER(c_int64), c_int]),
    "pc_group_by_tag": (c_int, [_P, _P, _P, _P, _P, _P, POINTER(GroupParams), c_int64, _P]),
    "pc_transform_keypoints": (
        c_int,
        [_P, _P, _P, _P, _P, c_float, c_int32, c_int32, c_int64, _P],
    ),
    "pc_refine_missing": (c_int, [_P, _P, _P, _P, _P, POINTER(RefineParams), c_int64, _P]),
    "pc_oks_nms": (c_int, [_P, _P, _P, _P, _P, _P, POINTER(OksNmsParams), c_int64, _P]),
    "pc_oks_nms_f64": (c_int, [_P, _P, _P, _P, _P, _P, POINTER(OksNmsParams), c_int64, _P]),
    "pc_scatter_results": (
        c_int, [_P, _P, POINTER(c_void_p), c_int32, _P, c_int64, c_int32, c_int64, _P]),
    "pc_scatter_results_signal": (
        c_int, [_P, _P, POINTER(c_void_p), c_int32, _P, c_int64, c_int32, c_int64,
                POINTER(c_void_p), c_int32, c_int32, _P, _P, _P]),
    "pc_wait_peer_flags": (c_int, [_P, c_int32, _P, ctypes.c_uint32, _P]),
    "pc_topdown_decode_gather": (
        c_int, [_P, _P, _P, _P, _P, _P, _P, POINTER(TopDownDecodeParams), c_int64,
                POINTER(GatherTarget), _P]),
    "pc_ctx_create": (c_int, [c_int, c_int64, POINTER(c_void_p)]),
    "pc_ctx_destroy": (c_int, [c_void_p]),
    "pc_ctx_last_transfer_bytes": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int64)]),
    "pc_crop_source_rect": (c_int, [_P, c_float, POINTER(AffineHostParams), _P]),
    "pc_topdown_affine_host": (
        c_int,
        [c_void_p, _P, _P, _P, _P, _P, _P, POINTER(AffineHostParams), c_int64],
    ),
    "pc_topdown_decode_host": (
        c_int,
        [c_void_p, _P, _P, _P, _P, _P, _P, _P, POINTER(TopDownDecodeParams), c_int64],
    ),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load libposecodec.so (once). Raises RuntimeError if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA codec was not built. "
            "Run `python -m mindpose_b200.csrc.build` (or __graft_entry__.build()). "
            "There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status: int) -> None:
    """Raise the reference-style exception for a negative pc_status."""
    if status == PC_OK:
        return
    msg = load().pc_last_error().decode("utf-8", "replace")
    if status in (PC_ERR_INVALID_ARGUMENT, PC_ERR_UNSUPPORTED):
        raise ValueError(msg)
    raise RuntimeError(f"libposecodec: {msg} (status {status})")


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args))


def device_ptr(obj) -> int:
    """Raw device address of a CUDA array (torch.Tensor, or any object exposing
    ``__cuda_array_interface__``); None -> NULL."""
    if obj is None:
        return 0
    if hasattr(obj, "data_ptr"):  # torch.Tensor
        if not obj.is_cuda:
            raise ValueError("expected a CUDA tensor; host tensors are not accepted here")
        if not obj.is_contiguous():
            raise ValueError("expected a contiguous tensor")
        return int(obj.data_ptr())
    cai = getattr(obj, "__cuda_array_interface__", None)
    if cai is not None:
        if cai.get("strides") is not None:
            raise ValueError("expected a contiguous CUDA array")
        return int(cai["data"][0])
    raise ValueError(f"cannot take a device pointer from {type(obj).__name__}")


def host_ptr(arr) -> int:
    """Address of a C-contiguous numpy array; None -> NULL."""
    if arr is None:
        return 0
    if not arr.flags["C_CONTIGUOUS"]:
        raise ValueError("expected a C-contiguous numpy array")
    return int(arr.ctypes.data)


def current_stream() -> int:
    import torch

    return int(torch.cuda.current_stream().cuda_stream)


def device_info(device: int = 0):
    sm, major, minor = c_int(0), c_int(0), c_int(0)
    call("pc_device_info", device, ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor))
    return sm.value, major.value, minor.value
