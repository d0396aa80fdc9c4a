"""Bottom-up codec entry points: fused decode, tag grouping, back-projection
(torch CUDA tensors; one libposecodec call each)."""
import ctypes
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib


def _same_storage_view(view: torch.Tensor, base: torch.Tensor, start_channel: int) -> bool:
    """True if `view` is `base[:, start_channel:start_channel + C]` of a contiguous base."""
    if not base.is_contiguous() or view.dim() != 4 or base.dim() != 4:
        return False
    plane = base.shape[2] * base.shape[3]
    return (view.data_ptr() == base.data_ptr() + start_channel * plane * base.element_size()
            and view.stride() == base.stride() and view.shape[0] == base.shape[0]
            and view.shape[2:] == base.shape[2:])


def decode(decoder, heatmap: List[torch.Tensor], tagging_heatmap: List[torch.Tensor],
           mask: torch.Tensor, raw: Optional[List[torch.Tensor]] = None):
    """``BottomUpHeatMapAEDecoder.decode``.  ``heatmap`` / ``tagging_heatmap`` are the
    lists ``decouple_output`` returns (channel views of the network outputs); the
    kernel reads the network outputs in place, so the views must come from
    ``decouple_output`` (or ``raw`` must hold the outputs they were cut from)."""
    k = decoder.num_joints
    stages = decoder.num_stages
    if stages not in (1, 2):
        raise ValueError("num_stages must be 1 or 2")
    if list(decoder.with_ae_loss[:stages]) != ([True, False][:stages]):
        raise ValueError("the CUDA decoder supports with_ae_loss=[True, False] (tags on the "
                         "first stage only, the HigherHRNet recipe)")
    per_joint = bool(decoder.tag_per_joint)
    c0 = 2 * k if per_joint else k + 1   # heat | tag planes of the first-stage output
    if len(heatmap) != stages or len(tagging_heatmap) != 1:
        raise ValueError("expected one heatmap per stage and one tag map")
    # recover the contiguous first-stage output the two views were sliced from; a candidate
    # (the caller's `raw[0]` or the view's base) is only used when both views are provably
    # its channel slices -- same pointer arithmetic, strides, N, H and W
    out0 = raw[0] if raw is not None else heatmap[0]._base
    if out0 is None or out0.dim() != 4 or out0.shape[1] != c0 \
            or not _same_storage_view(heatmap[0], out0, 0) \
            or not _same_storage_view(tagging_heatmap[0], out0, k):
        # views of something else: pack heat | tag into one contiguous tensor
        out0 = torch.cat([heatmap[0], tagging_heatmap[0]], dim=1).contiguous()
        if out0.shape[1] != c0:
            raise ValueError(f"heatmap[0] | tagging_heatmap[0] must hold {c0} channels "
                             f"(tag_per_joint={per_joint}), got {out0.shape[1]}")
    out1 = None
    if stages == 2:
        # the second stage is one tensor of K channels: the view itself is what is decoded
        # (a batch-sliced view has a contiguous base of the same K channels but other images,
        # so the base is never substituted for it)
        out1 = heatmap[1]
        if raw is not None and _same_storage_view(out1, raw[1], 0) and raw[1].shape[1] == k:
            out1 = raw[1]
        if out1.dim() != 4 or out1.shape[1] != k:
            raise ValueError(f"heatmap[1] must be [N, {k}, H, W]")
        if not out1.is_contiguous():
            out1 = out1.contiguous()
        if out1.shape[0] != out0.shape[0]:
            raise ValueError("heatmap[0] and heatmap[1] hold different numbers of images")
    for t in (out0, out1):
        if t is not None and (not t.is_cuda or t.dtype != torch.float32):
            raise ValueError("network outputs must be float32 CUDA tensors")
    n = out0.shape[0]
    if stages == 2:
        h0, w0 = out0.shape[2:]
        h1, w1 = out1.shape[2:]
    else:
        h0 = w0 = 0
        h1, w1 = out0.shape[2:]
    if not mask.is_cuda:
        raise ValueError("`mask` must live on a CUDA device")
    mask_u8 = (mask != 0).to(torch.uint8).contiguous() if mask.dtype != torch.uint8 else mask.contiguous()
    mh, mw = mask_u8.shape[1:]
    m = decoder.max_num
    dev = out0.device
    val_k = torch.empty((n, k, m), dtype=torch.float32, device=dev)
    tag_k = torch.empty((n, k, m, 1), dtype=torch.float32, device=dev)
    ind_k = torch.empty((n, k, m, 2), dtype=torch.float32, device=dev)
    raw_map = tag_map = None
    if getattr(decoder, "return_maps", True):
        raw_map = torch.empty((n, k, h1, w1), dtype=torch.float32, device=dev)
        tag_map = torch.empty((n, k if per_joint else 1, h1, w1, 1), dtype=torch.float32,
                              device=dev)
    p = _lib.BottomUpDecodeParams(k, stages, h0, w0, h1, w1, mh, mw, int(bool(decoder.use_nms)),
                                  int(decoder.nms_kernel), m, int(bool(decoder.shift_coordinate)),
                                  int(per_joint))
    with torch.cuda.device(dev):
        _lib.call("pc_bottomup_decode", _lib.device_ptr(out0), _lib.device_ptr(out1),
                  _lib.device_ptr(mask_u8), _lib.device_ptr(val_k), _lib.device_ptr(tag_k),
                  _lib.device_ptr(ind_k), _lib.device_ptr(raw_map), _lib.device_ptr(tag_map),
                  ctypes.byref(p), n, _lib.current_stream())
    return val_k, tag_k, ind_k, raw_map, tag_map


def decode_stats(reset: bool = False) -> int:
    """Planes (image x joint) that needed the exact second pass since the last reset
    (``pc_bottomup_decode_stats``; synchronises the device)."""
    v = ctypes.c_int64(0)
    _lib.call("pc_bottomup_decode_stats", ctypes.byref(v), int(reset))
    return int(v.value)


def max_group_capacity(num_joints: int, max_num: int) -> int:
    """The largest ``max_groups`` ``group_by_tag`` accepts: what 227 KB of shared memory holds
    (csrc/grouping.cu: 44 + 4 * (D + K) bytes per group, D = 32 or 64 detections per joint),
    and never more than K * max_num -- every detection a person of its own."""
    d = 32 if max_num <= 32 else 64
    room = (227 * 1024 - 32 * d) // (44 + 4 * (d + num_joints))
    return min(num_joints * max_num, room)


def group_by_tag(val_k: torch.Tensor, tag_k: torch.Tensor, ind_k: torch.Tensor,
                 joint_order: Sequence[int], vis_thr: float = 0.1, tag_thr: float = 1.0,
                 ignore_too_much: bool = False, use_rounded_norm: bool = True,
                 max_groups: Optional[int] = None):
    """Batched ``match_by_tag`` + instance score.

    -> (ans f32 [N, G, K, 4], num_groups i32 [N], scores f32 [N, G]) with G = ``max_groups``
    (default PC_MAX_GROUPS = 128); image i has ``num_groups[i]`` people in
    ``ans[i, :num_groups[i]]`` (insertion order); -1 flags more people than G.  The reference
    is unbounded; ``max_groups = K * M`` (every detection its own group) can never overflow."""
    for t in (val_k, tag_k, ind_k):
        if not (t.is_cuda and t.dtype == torch.float32):
            raise ValueError("val_k / tag_k / ind_k must be float32 CUDA tensors")
    n, k, m = val_k.shape
    if tag_k.shape != (n, k, m, 1):
        raise ValueError("the CUDA grouping supports one tag channel (tag_k [N,K,M,1])")
    val_k, tag_k, ind_k = val_k.contiguous(), tag_k.contiguous(), ind_k.contiguous()
    dev = val_k.device
    g = _lib.PC_MAX_GROUPS if max_groups is None else int(max_groups)
    if g < 1:
        raise ValueError("`max_groups` must be >= 1")
    ans = torch.empty((n, g, k, 4), dtype=torch.float32, device=dev)
    num = torch.empty((n,), dtype=torch.int32, device=dev)
    scores = torch.zeros((n, g), dtype=torch.float32, device=dev)
    p = _lib.GroupParams()
    p.num_joints, p.max_num = k, m
    p.vis_thr, p.tag_thr = float(vis_thr), float(tag_thr)
    p.ignore_too_much, p.use_rounded_norm = int(bool(ignore_too_much)), int(bool(use_rounded_norm))
    p.max_groups = g
    order = list(joint_order)
    if len(order) != k:
        raise ValueError("`joint_order` must list every joint once")
    for i, j in enumerate(order):
        p.joint_order[i] = int(j)
    with torch.cuda.device(dev):
        _lib.call("pc_group_by_tag", _lib.device_ptr(val_k), _lib.device_ptr(tag_k),
                  _lib.device_ptr(ind_k), _lib.device_ptr(ans), _lib.device_ptr(num),
                  _lib.device_ptr(scores), ctypes.byref(p), n, _lib.current_stream())
    return ans, num, scores


def transform_keypoints(ans: torch.Tensor, num_groups: torch.Tensor, center, scale, heatmap_wh,
                        pixel_std: float = 200.0) -> torch.Tensor:
    """In-place back-projection of grouped people (utils.py:235-274). center / scale /
    heatmap_wh: [N,2] (any float dtype; used as float64 like the reference's numpy)."""
    dev = ans.device

    def f64(x):
        return torch.as_tensor(np.asarray(x.cpu() if isinstance(x, torch.Tensor) else x),
                               dtype=torch.float64).reshape(-1, 2).to(dev).contiguous()

    n, g, k, _ = ans.shape
    c, s, hw = f64(center), f64(scale), f64(heatmap_wh)
    with torch.cuda.device(dev):
        _lib.call("pc_transform_keypoints", _lib.device_ptr(ans), _lib.device_ptr(num_groups),
                  _lib.device_ptr(c), _lib.device_ptr(s), _lib.device_ptr(hw), float(pixel_std),
                  k, g, n, _lib.current_stream())
    return ans


def encode_targets(keypoints: torch.Tensor, heatmap_sizes, sigma: float = 2.0, max_num: int = 30,
                   tag_per_joint: bool = True):
    """Batched ``BottomUpGenerateTarget._encoding`` (bottomup_transform.py:504-598).

    keypoints f32 [N, S, M, K, 3]: per scale, the M people's joints in heat-map pixels of that
    scale (pad people with visibility 0). heatmap_sizes = [[w, h], ...] (S entries).
    -> (target f32 [N, S, K, Hmax, Wmax], tag_ind i32 [N, S, max_num, K, 2] or [N, S, max_num, 2])."""
    if not (isinstance(keypoints, torch.Tensor) and keypoints.is_cuda):
        raise ValueError("`keypoints` must be a CUDA tensor (no CPU fallback)")
    if keypoints.dim() != 5 or keypoints.shape[-1] != 3:
        raise ValueError("`keypoints` must have shape [N, S, M, K, 3]")
    keypoints = keypoints.to(torch.float32).contiguous()
    n, s, m, k, _ = keypoints.shape
    sizes = np.asarray(heatmap_sizes).reshape(-1, 2)
    if sizes.shape[0] != s:
        raise ValueError("one heatmap size per scale is required")
    if s > _lib.PC_MAX_SCALES:
        raise ValueError(f"at most {_lib.PC_MAX_SCALES} scales are supported")
    p = _lib.BottomUpEncodeParams()
    p.num_joints, p.num_scales, p.num_people, p.max_num = k, s, m, int(max_num)
    for i in range(s):
        p.heatmap_w[i], p.heatmap_h[i] = int(sizes[i, 0]), int(sizes[i, 1])
    p.sigma = float(sigma)
    p.tag_per_joint = int(bool(tag_per_joint))
    hmax, wmax = int(sizes[:, 1].max()), int(sizes[:, 0].max())
    dev = keypoints.device
    target = torch.empty((n, s, k, hmax, wmax), dtype=torch.float32, device=dev)
    tag_shape = (n, s, int(max_num), k, 2) if tag_per_joint else (n, s, int(max_num), 2)
    tag_ind = torch.empty(tag_shape, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("pc_bottomup_encode", _lib.device_ptr(keypoints) if m else 0,
                  _lib.device_ptr(target), _lib.device_ptr(tag_ind), ctypes.byref(p), n,
                  _lib.current_stream())
    return target, tag_ind


def refine_missing(heatmap: torch.Tensor, tagging_heatmap: torch.Tensor, ans: torch.Tensor,
                   num_groups: torch.Tensor) -> torch.Tensor:
    """Batched ``_refine_missing`` (bottomup_inferencer.py:189-249), in place on ``ans``
    [N, G, K, 4] (heat-map coordinates; G as ``group_by_tag`` made it).  heatmap [N,K,H,W] / tagging_heatmap
    [N,K,H,W,1] are the decoder's ``heatmap_raw`` / ``tagging_heatmap`` outputs."""
    for t in (heatmap, tagging_heatmap, ans):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError("heatmap / tagging_heatmap / ans must be contiguous float32 CUDA tensors")
    n, k, h, w = heatmap.shape
    if tagging_heatmap.shape != (n, k, h, w, 1):
        raise ValueError("the CUDA refinement supports one tag channel (tagging_heatmap [N,K,H,W,1])")
    if ans.dim() != 4 or ans.shape[0] != n or ans.shape[2:] != (k, 4):
        raise ValueError("`ans` must be the [N, G, K, 4] output of group_by_tag")
    g = ans.shape[1]
    num_groups = num_groups.to(torch.int32).contiguous()
    scratch = torch.empty((n, g), dtype=torch.float32, device=ans.device)
    p = _lib.RefineParams(k, h, w, g)
    with torch.cuda.device(ans.device):
        _lib.call("pc_refine_missing", _lib.device_ptr(heatmap), _lib.device_ptr(tagging_heatmap),
                  _lib.device_ptr(ans), _lib.device_ptr(num_groups), _lib.device_ptr(scratch),
                  ctypes.byref(p), n, _lib.current_stream())
    return ans
