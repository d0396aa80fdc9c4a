"""Functional, batched entry points of the heatmap codec (torch CUDA tensors in,
torch CUDA tensors out).  Each function is one call into libposecodec through
``_lib`` on the current torch stream; torch is used for device memory and
streams only.  Host (numpy) inputs are accepted only by the ``*_host``
functions, which go through the library's own pipelined copy front end.
"""
import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise ValueError(f"`{name}` must be a torch.Tensor on a CUDA device")
    if not t.is_cuda:
        raise ValueError(f"`{name}` must live on a CUDA device (no CPU fallback)")
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()


def _wh(size) -> Tuple[int, int]:
    w, h = (int(v) for v in np.asarray(size).reshape(-1)[:2])
    return w, h


# --------------------------------------------------------------------------- A1
def box_to_center_scale(boxes: torch.Tensor, image_size, pixel_std: float = 200.0,
                        scale_padding: float = 1.25):
    """boxes f32 [N,4] (x,y,w,h) -> (center f32 [N,2], scale f32 [N,2])."""
    boxes = _f32(boxes, "boxes")
    n = boxes.shape[0]
    center = torch.empty((n, 2), dtype=torch.float32, device=boxes.device)
    scale = torch.empty((n, 2), dtype=torch.float32, device=boxes.device)
    w, h = _wh(image_size)
    p = _lib.BoxParams(w, h, float(pixel_std), float(scale_padding))
    with torch.cuda.device(boxes.device):
        _lib.call("pc_box_to_center_scale", _lib.device_ptr(boxes), _lib.device_ptr(center),
                  _lib.device_ptr(scale), ctypes.byref(p), n, _lib.current_stream())
    return center, scale


# ----------------------------------------------------------------------- A2 / A3
def affine_matrices(center: torch.Tensor, scale: torch.Tensor, rot: Optional[torch.Tensor],
                    image_size, pixel_std: float = 200.0, use_udp: bool = False):
    """-> (fwd f64 [N,2,3], inv f64 [N,2,3])."""
    center = _f32(center, "center")
    scale = _f32(scale, "scale")
    n = center.shape[0]
    if rot is not None:
        rot = _f32(rot, "rot").reshape(n)
    fwd = torch.empty((n, 2, 3), dtype=torch.float64, device=center.device)
    inv = torch.empty((n, 2, 3), dtype=torch.float64, device=center.device)
    w, h = _wh(image_size)
    p = _lib.AffineParams(w, h, float(pixel_std), int(bool(use_udp)))
    with torch.cuda.device(center.device):
        _lib.call("pc_affine_matrices", _lib.device_ptr(center), _lib.device_ptr(scale),
                  _lib.device_ptr(rot), _lib.device_ptr(fwd), _lib.device_ptr(inv),
                  ctypes.byref(p), n, _lib.current_stream())
    return fwd, inv


def affine_from_points(src_points: torch.Tensor, dst_points: torch.Tensor):
    """``cv2.getAffineTransform`` for caller-built float32 point triples [N,3,2]
    -> (fwd f64 [N,2,3], inv f64 [N,2,3])."""
    src_points = _f32(src_points, "src_points").reshape(-1, 3, 2)
    dst_points = _f32(dst_points, "dst_points").reshape(-1, 3, 2)
    n = src_points.shape[0]
    if dst_points.shape[0] != n:
        raise ValueError("`src_points` and `dst_points` must hold the same number of triples")
    fwd = torch.empty((n, 2, 3), dtype=torch.float64, device=src_points.device)
    inv = torch.empty((n, 2, 3), dtype=torch.float64, device=src_points.device)
    with torch.cuda.device(src_points.device):
        _lib.call("pc_affine_from_points", _lib.device_ptr(src_points),
                  _lib.device_ptr(dst_points), _lib.device_ptr(fwd), _lib.device_ptr(inv), n,
                  _lib.current_stream())
    return fwd, inv


def invert_affine(fwd: torch.Tensor) -> torch.Tensor:
    if not fwd.is_cuda:
        raise ValueError("`fwd` must live on a CUDA device")
    fwd = fwd.to(torch.float64).contiguous().reshape(-1, 2, 3)
    inv = torch.empty_like(fwd)
    with torch.cuda.device(fwd.device):
        _lib.call("pc_invert_affine", _lib.device_ptr(fwd), _lib.device_ptr(inv), fwd.shape[0],
                  _lib.current_stream())
    return inv


# --------------------------------------------------------------------------- A4
def warp_affine(src: torch.Tensor, src_offset: torch.Tensor, src_hw: torch.Tensor,
                inv: torch.Tensor, dst_size, channels: int = 3,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """src: u8 buffer holding dense HWC images; crop i reads the image at byte
    offset src_offset[i] (i64) of size src_hw[i] = (rows, cols) (i32).
    inv f64 [N,2,3]; dst_size = [w, h]. -> u8 [N, h, w, C]."""
    if not (src.is_cuda and src.dtype == torch.uint8 and src.is_contiguous()):
        raise ValueError("`src` must be a contiguous uint8 CUDA tensor")
    n = inv.shape[0]
    src_offset = src_offset.to(torch.int64).contiguous()
    src_hw = src_hw.to(torch.int32).contiguous()
    inv = inv.to(torch.float64).contiguous()
    w, h = _wh(dst_size)
    if out is None:
        out = torch.empty((n, h, w, channels), dtype=torch.uint8, device=src.device)
    p = _lib.WarpParams(w, h, int(channels))
    with torch.cuda.device(src.device):
        _lib.call("pc_warp_affine_u8", _lib.device_ptr(src), _lib.device_ptr(src_offset),
                  _lib.device_ptr(src_hw), _lib.device_ptr(inv), _lib.device_ptr(out),
                  ctypes.byref(p), n, _lib.current_stream())
    return out


def warp_affine_normalized(src: torch.Tensor, src_offset: torch.Tensor, src_hw: torch.Tensor,
                           inv: torch.Tensor, dst_size, mean: Sequence[float],
                           std: Sequence[float], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Warp + ``vision.Normalize(mean, std)`` + ``HWC2CHW`` in one pass
    (mindpose/data/data_factory.py:127-138): -> f32 [N, 3, h, w].  ``mean`` / ``std`` are the
    values given to Normalize, i.e. already multiplied by 255."""
    if not (src.is_cuda and src.dtype == torch.uint8 and src.is_contiguous()):
        raise ValueError("`src` must be a contiguous uint8 CUDA tensor")
    n = inv.shape[0]
    src_offset = src_offset.to(torch.int64).contiguous()
    src_hw = src_hw.to(torch.int32).contiguous()
    inv = inv.to(torch.float64).contiguous()
    w, h = _wh(dst_size)
    mean = np.asarray(mean, dtype=np.float32).reshape(-1)
    std = np.asarray(std, dtype=np.float32).reshape(-1)
    if mean.shape[0] != 3 or std.shape[0] != 3:
        raise ValueError("`mean` and `std` must have three entries")
    if out is None:
        out = torch.empty((n, 3, h, w), dtype=torch.float32, device=src.device)
    p = _lib.WarpNormParams()
    p.dst_w, p.dst_h, p.channels = w, h, 3
    for c in range(3):
        p.mean[c], p.std[c] = float(mean[c]), float(std[c])
    with torch.cuda.device(src.device):
        _lib.call("pc_warp_affine_u8_norm_chw", _lib.device_ptr(src), _lib.device_ptr(src_offset),
                  _lib.device_ptr(src_hw), _lib.device_ptr(inv), _lib.device_ptr(out),
                  ctypes.byref(p), n, _lib.current_stream())
    return out


def warp_affine_uniform(images: torch.Tensor, inv: torch.Tensor, dst_size) -> torch.Tensor:
    """images u8 [N, Hs, Ws, C], one source image per crop."""
    n, hs, ws, c = images.shape
    off = torch.arange(n, device=images.device, dtype=torch.int64) * (hs * ws * c)
    hw = torch.tensor([hs, ws], device=images.device, dtype=torch.int32).repeat(n, 1)
    return warp_affine(images, off, hw, inv, dst_size, channels=c)


def rescale_pad(images: torch.Tensor, src_offset: torch.Tensor, src_hw: torch.Tensor,
                dst_wh: torch.Tensor, canvas_wh, with_mask: bool = True,
                mean: Optional[Sequence[float]] = None, std: Optional[Sequence[float]] = None):
    """Bilinear rescale (cv2.resize INTER_LINEAR arithmetic) of every image to its own target
    size into the top-left corner of a zero canvas, plus the validity mask.

    images u8 (any shape: one allocation holding N HWC images), src_offset i64 [N] (bytes),
    src_hw i32 [N,2] (height, width), dst_wh i32 [N,2] (width, height)
    -> (canvas u8 [N, canvas_h, canvas_w, 3], mask u8 [N, canvas_h, canvas_w] or None).

    With ``mean`` / ``std`` (the values given to ``vision.Normalize``, i.e. already multiplied
    by 255) the pipeline's Normalize + HWC2CHW step (data_factory.py:127-138) is fused in and
    the canvas comes back as float32 [N, 3, canvas_h, canvas_w]."""
    if not (images.is_cuda and images.dtype == torch.uint8 and images.is_contiguous()):
        raise ValueError("`images` must be a contiguous uint8 CUDA tensor")
    cw, ch = _wh(canvas_wh)
    n = src_offset.shape[0]
    dev = images.device
    src_offset = src_offset.to(device=dev, dtype=torch.int64).contiguous()
    src_hw = src_hw.to(device=dev, dtype=torch.int32).contiguous()
    dst_wh = dst_wh.to(device=dev, dtype=torch.int32).contiguous()
    if src_hw.shape != (n, 2) or dst_wh.shape != (n, 2):
        raise ValueError("`src_hw` and `dst_wh` must be [N, 2]")
    mask = torch.empty((n, ch, cw), dtype=torch.uint8, device=dev) if with_mask else None
    if (mean is None) != (std is None):
        raise ValueError("give both `mean` and `std`, or neither")
    if mean is not None:
        mean = np.asarray(mean, dtype=np.float32).reshape(-1)
        std = np.asarray(std, dtype=np.float32).reshape(-1)
        if mean.shape[0] != 3 or std.shape[0] != 3:
            raise ValueError("`mean` and `std` must have three entries")
        out = torch.empty((n, 3, ch, cw), dtype=torch.float32, device=dev)
        p = _lib.WarpNormParams()
        p.dst_w, p.dst_h, p.channels = cw, ch, 3
        for c in range(3):
            p.mean[c], p.std[c] = float(mean[c]), float(std[c])
        with torch.cuda.device(dev):
            _lib.call("pc_rescale_pad_u8_norm_chw", _lib.device_ptr(images),
                      _lib.device_ptr(src_offset), _lib.device_ptr(src_hw), _lib.device_ptr(dst_wh),
                      _lib.device_ptr(out), _lib.device_ptr(mask), ctypes.byref(p), n,
                      _lib.current_stream())
        return out, mask
    out = torch.empty((n, ch, cw, 3), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.call("pc_rescale_pad_u8", _lib.device_ptr(images), _lib.device_ptr(src_offset),
                  _lib.device_ptr(src_hw), _lib.device_ptr(dst_wh), _lib.device_ptr(out),
                  _lib.device_ptr(mask), cw, ch, 3, n, _lib.current_stream())
    return out, mask


def affine_joints(keypoints: torch.Tensor, fwd: torch.Tensor, use_udp: bool = False):
    """In place on keypoints f32 [N,K,3]; returns it."""
    if not (keypoints.is_cuda and keypoints.dtype == torch.float32 and keypoints.is_contiguous()):
        raise ValueError("`keypoints` must be a contiguous float32 CUDA tensor")
    n, k = keypoints.shape[:2]
    fwd = fwd.to(torch.float64).contiguous()
    with torch.cuda.device(keypoints.device):
        _lib.call("pc_affine_joints", _lib.device_ptr(keypoints), _lib.device_ptr(fwd), k,
                  int(bool(use_udp)), n, _lib.current_stream())
    return keypoints


# ----------------------------------------------------------------------- A5 / A6
def topdown_encode(keypoints: torch.Tensor, image_size, heatmap_size, sigma: float = 2.0,
                   use_udp: bool = False, joint_weights: Optional[Sequence[float]] = None,
                   out: Optional[torch.Tensor] = None):
    """keypoints f32 [N,K,3] -> (target f32 [N,K,H,W], target_weight f32 [N,K])."""
    keypoints = _f32(keypoints, "keypoints")
    n, k = keypoints.shape[:2]
    iw, ih = _wh(image_size)
    w, h = _wh(heatmap_size)
    p = _lib.EncodeParams()
    p.num_joints, p.image_w, p.image_h, p.heatmap_w, p.heatmap_h = k, iw, ih, w, h
    p.sigma = float(sigma)
    p.use_udp = int(bool(use_udp))
    p.use_joint_weights = 0
    if joint_weights is not None:
        jw = np.asarray(joint_weights, dtype=np.float32).reshape(-1)
        if jw.shape[0] != k:
            raise ValueError("`joint_weights` must have one entry per joint")
        p.use_joint_weights = 1
        for i in range(k):
            p.joint_weights[i] = float(jw[i])
    target = out if out is not None else torch.empty(
        (n, k, h, w), dtype=torch.float32, device=keypoints.device)
    weight = torch.empty((n, k), dtype=torch.float32, device=keypoints.device)
    with torch.cuda.device(keypoints.device):
        _lib.call("pc_topdown_encode", _lib.device_ptr(keypoints), _lib.device_ptr(target),
                  _lib.device_ptr(weight), ctypes.byref(p), n, _lib.current_stream())
    return target, weight


# ---------------------------------------------------------------------- A8 - A13
def make_decode_params(num_joints: int, height: int, width: int, pixel_std: float = 200.0,
                       to_original: bool = True, shift_coordinate: bool = False,
                       use_udp: bool = False, dark_udp_refine: bool = False,
                       kernel_size: int = 11, flip_index=None, shift_heatmap: bool = False,
                       dark_kernel: Optional[np.ndarray] = None) -> "_lib.TopDownDecodeParams":
    p = _lib.TopDownDecodeParams()
    p.num_joints, p.height, p.width = int(num_joints), int(height), int(width)
    p.pixel_std = float(pixel_std)
    p.to_original = int(bool(to_original))
    p.shift_coordinate = int(bool(shift_coordinate))
    p.use_udp = int(bool(use_udp))
    p.dark_udp_refine = int(bool(dark_udp_refine))
    p.kernel_size = int(kernel_size)
    p.flip_test = 0
    p.shift_heatmap = int(bool(shift_heatmap))
    if flip_index is not None:
        fi = np.asarray(flip_index).reshape(-1)
        if fi.shape[0] != num_joints:
            raise ValueError("`flip_index` must have one entry per joint")
        p.flip_test = 1
        for i in range(num_joints):
            p.flip_index[i] = int(fi[i])
    p.dark_kernel_set = 0
    if dark_kernel is not None:
        dk = np.asarray(dark_kernel, dtype=np.float32).reshape(-1)
        if dk.shape[0] != kernel_size * kernel_size or kernel_size > _lib.PC_MAX_DARK_KERNEL:
            raise ValueError("`dark_kernel` must hold kernel_size^2 floats, kernel_size <= 17")
        p.dark_kernel_set = 1
        for i in range(dk.shape[0]):
            p.dark_kernel[i] = float(dk[i])
    return p


def topdown_decode(heatmap: torch.Tensor, center: torch.Tensor, scale: torch.Tensor,
                   score: torch.Tensor, flipped: Optional[torch.Tensor] = None,
                   params: Optional["_lib.TopDownDecodeParams"] = None,
                   out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, gather=None,
                   **kwargs):
    """heatmap f32 [N,K,H,W] (+ optional flipped pair) -> (all_preds [N,K,3], all_boxes [N,6]);
    ``out`` = (preds, boxes) writes into caller-owned contiguous float32 tensors.

    ``gather`` (a ``dist.PeerGather``): the decode kernel also stores every result into the
    gathered table of every rank and signals the step (``pc_topdown_decode_gather``: decode and
    all-gather as one kernel); wait for the other ranks' rows with ``gather.wait_lag`` /
    the ticket ``gather.last_ticket()``."""
    heatmap = _f32(heatmap, "heatmap")
    if heatmap.dim() != 4:
        raise ValueError("`heatmap` must have shape [N, K, H, W]")
    n, k, h, w = heatmap.shape
    if params is None:
        params = make_decode_params(k, h, w, **kwargs)
    center = _f32(center, "center").reshape(n, 2)
    scale = _f32(scale, "scale").reshape(n, 2)
    score = _f32(score, "score").reshape(n)
    if params.flip_test:
        if flipped is None:
            raise ValueError("flip test needs the flipped heatmap")
        flipped = _f32(flipped, "flipped")
        if flipped.shape != heatmap.shape:
            raise ValueError("`flipped` must have the shape of `heatmap`")
    else:
        flipped = None
    if out is not None:
        preds, boxes = out
        for t, shape in ((preds, (n, k, 3)), (boxes, (n, 6))):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
                    and tuple(t.shape) == shape):
                raise ValueError(f"`out` must hold contiguous float32 CUDA tensors [N,K,3], [N,6]; "
                                 f"got {tuple(t.shape)} for {shape}")
    else:
        preds = torch.empty((n, k, 3), dtype=torch.float32, device=heatmap.device)
        boxes = torch.empty((n, 6), dtype=torch.float32, device=heatmap.device)
    with torch.cuda.device(heatmap.device):
        if gather is not None:
            target = gather.next_target(n, k)      # bumps the gather's step
            _lib.call("pc_topdown_decode_gather", _lib.device_ptr(heatmap),
                      _lib.device_ptr(flipped), _lib.device_ptr(center), _lib.device_ptr(scale),
                      _lib.device_ptr(score), _lib.device_ptr(preds), _lib.device_ptr(boxes),
                      ctypes.byref(params), n, ctypes.byref(target), _lib.current_stream())
        else:
            _lib.call("pc_topdown_decode", _lib.device_ptr(heatmap), _lib.device_ptr(flipped),
                      _lib.device_ptr(center), _lib.device_ptr(scale), _lib.device_ptr(score),
                      _lib.device_ptr(preds), _lib.device_ptr(boxes), ctypes.byref(params), n,
                      _lib.current_stream())
    return preds, boxes


_UPLOAD_MODES = {"full": _lib.UPLOAD_FULL, "roi": _lib.UPLOAD_ROI,
                 "roi_kernel": _lib.UPLOAD_ROI_KERNEL}
DEFAULT_UPLOAD = "roi_kernel"   # measured: profiles/README.md (r01r / r01s / r01t)


class HostContext:
    """Owns a pc_ctx: device scratch + two streams for the host-buffer path."""

    def __init__(self, device: int = 0, scratch_bytes: int = 1 << 30):
        self._h = ctypes.c_void_p()
        _lib.call("pc_ctx_create", int(device), int(scratch_bytes), ctypes.byref(self._h))

    def close(self):
        if self._h:
            _lib.load().pc_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def topdown_decode(self, heatmap: np.ndarray, center: np.ndarray, scale: np.ndarray,
                       score: np.ndarray, flipped: Optional[np.ndarray] = None,
                       params: Optional["_lib.TopDownDecodeParams"] = None,
                       out_preds: Optional[np.ndarray] = None,
                       out_boxes: Optional[np.ndarray] = None, **kwargs):
        """numpy (host, ideally pinned) in -> numpy out; copies are pipelined
        with the kernel inside the library."""
        heatmap = np.ascontiguousarray(heatmap, dtype=np.float32)
        n, k, h, w = heatmap.shape
        if params is None:
            params = make_decode_params(k, h, w, **kwargs)
        center = np.ascontiguousarray(center, dtype=np.float32).reshape(n, 2)
        scale = np.ascontiguousarray(scale, dtype=np.float32).reshape(n, 2)
        score = np.ascontiguousarray(score, dtype=np.float32).reshape(n)
        if params.flip_test:
            if flipped is None:
                raise ValueError("flip test needs the flipped heatmap")
            flipped = np.ascontiguousarray(flipped, dtype=np.float32)
        else:
            flipped = None
        preds = out_preds if out_preds is not None else np.empty((n, k, 3), np.float32)
        boxes = out_boxes if out_boxes is not None else np.empty((n, 6), np.float32)
        _lib.call("pc_topdown_decode_host", self._h, _lib.host_ptr(heatmap),
                  _lib.host_ptr(flipped), _lib.host_ptr(center), _lib.host_ptr(scale),
                  _lib.host_ptr(score), _lib.host_ptr(preds), _lib.host_ptr(boxes),
                  ctypes.byref(params), n)
        return preds, boxes

    def last_transfer_bytes(self):
        """(host->device, device->host) bytes the last call on this context moved."""
        h2d, d2h = ctypes.c_int64(0), ctypes.c_int64(0)
        _lib.call("pc_ctx_last_transfer_bytes", self._h, ctypes.byref(h2d), ctypes.byref(d2h))
        return int(h2d.value), int(d2h.value)

    def topdown_affine(self, images: np.ndarray, boxes: np.ndarray, image_size,
                       rot: Optional[np.ndarray] = None, pixel_std: float = 200.0,
                       scale_padding: float = 1.25, use_udp: bool = False,
                       out: Optional[np.ndarray] = None, upload: Optional[str] = None):
        """images u8 [N,Hs,Ws,C] + boxes f32 [N,4] (host) -> (crops u8 [N,h,w,C], center, scale).

        ``upload``: "full" copies whole source images to the device; "roi" copies only the
        rectangle each crop samples (one strided copy per crop); "roi_kernel" lets one kernel
        per chunk fetch the rectangles from PINNED ``images`` (falls back to "roi" for pageable
        memory).  Same crops in every mode, fewer PCIe bytes in the last two."""
        upload = DEFAULT_UPLOAD if upload is None else upload
        if upload not in _UPLOAD_MODES:
            raise ValueError(f"`upload` must be one of {sorted(_UPLOAD_MODES)}, got {upload!r}")
        if images.dtype != np.uint8 or images.ndim != 4:
            raise ValueError("`images` must be uint8 [N, Hs, Ws, C]")
        images = np.ascontiguousarray(images)
        n, hs, ws, c = images.shape
        boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(n, 4)
        if rot is not None:
            rot = np.ascontiguousarray(rot, dtype=np.float32).reshape(n)
        w, h = (int(v) for v in np.asarray(image_size).reshape(-1)[:2])
        crops = out if out is not None else np.empty((n, h, w, c), np.uint8)
        center = np.empty((n, 2), np.float32)
        scale = np.empty((n, 2), np.float32)
        p = _lib.AffineHostParams(hs, ws, c, w, h, float(pixel_std), float(scale_padding),
                                  int(bool(use_udp)), _UPLOAD_MODES[upload])
        _lib.call("pc_topdown_affine_host", self._h, _lib.host_ptr(images), _lib.host_ptr(boxes),
                  _lib.host_ptr(rot), _lib.host_ptr(crops), _lib.host_ptr(center),
                  _lib.host_ptr(scale), ctypes.byref(p), n)
        return crops, center, scale
