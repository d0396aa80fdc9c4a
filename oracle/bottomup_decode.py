"""Oracle (test infrastructure): HigherHRNet bottom-up heatmap decoding.

float32 numpy restatement of ``BottomUpHeatMapAEDecoder``
(mindpose/models/decoders/bottom_up_decoder.py:67-203), a MindSpore graph that
cannot be executed here.

PARITY UNPINNED.  The only reference test for this path is stale and asserts
shapes (tests/models/decoders/test_bottom_up_decoder.py:8-19).  MindSpore 1.x
semantics this restatement assumes:

* ``ops.ResizeBilinear(size)`` (align_corners=False) is the legacy asymmetric
  mapping ``src = dst * (in / out)`` (no half-pixel offset), ``lo = floor(src)``,
  ``hi = min(lo + 1, in - 1)``, and the lerp is
  ``top = tl + (tr - tl) * fx; bot = bl + (br - bl) * fx; top + (bot - top) * fy``
  in float32;
* ``_aggregate_heatmap`` (:129-138): ``(stage1 + resize(stage0)) / num_stages``;
* ``ops.ResizeNearestNeighbor`` (align_corners=False): ``src = min(floor(dst *
  in / out), in - 1)``; the resized float mask is cast to bool (non-zero = valid)
  and invalid pixels are filled with 0 (:123-125);
* ``nn.MaxPool2d(kernel_size=k, pad_mode="same")`` is stride 1 with implicit
  -inf padding (``(k-1)//2`` before, the rest after); ``_nms`` keeps ``h`` where
  ``h == pooled`` and multiplies the rest by 0 (:173-178);
* ``ops.top_k`` is sorted by value descending; among equal values the lowest
  flat index comes first (canonical rule of SURVEY.md note 3);
* tags are gathered from the bilinearly resized tag map at the top-k indices
  (:156-164); ``x = ind % W``, ``y = ind // W`` as float32 (:166-170);
* ``_shift_coordinate`` (:180-203) is restated WITH its quirk: the offsets come
  out of ``masked_select`` in row-major spatial order and are added to the
  coordinates in top-k rank order.
"""
import numpy as np

F32 = np.float32


def resize_bilinear_legacy(x, out_h, out_w):
    """x f32 [..., H, W] -> [..., out_h, out_w]."""
    x = np.asarray(x, dtype=F32)
    in_h, in_w = x.shape[-2:]
    sy = F32(in_h) / F32(out_h)
    sx = F32(in_w) / F32(out_w)
    ys = (np.arange(out_h, dtype=F32) * sy).astype(F32)
    xs = (np.arange(out_w, dtype=F32) * sx).astype(F32)
    y0 = np.floor(ys).astype(np.int64)
    x0 = np.floor(xs).astype(np.int64)
    y1 = np.minimum(y0 + 1, in_h - 1)
    x1 = np.minimum(x0 + 1, in_w - 1)
    fy = (ys - y0.astype(F32)).astype(F32)[:, None]
    fx = (xs - x0.astype(F32)).astype(F32)[None, :]
    tl = x[..., y0[:, None], x0[None, :]]
    tr = x[..., y0[:, None], x1[None, :]]
    bl = x[..., y1[:, None], x0[None, :]]
    br = x[..., y1[:, None], x1[None, :]]
    top = (tl + ((tr - tl).astype(F32) * fx).astype(F32)).astype(F32)
    bot = (bl + ((br - bl).astype(F32) * fx).astype(F32)).astype(F32)
    return (top + ((bot - top).astype(F32) * fy).astype(F32)).astype(F32)


def resize_nearest(x, out_h, out_w):
    in_h, in_w = x.shape[-2:]
    sy = F32(in_h) / F32(out_h)
    sx = F32(in_w) / F32(out_w)
    yi = np.minimum(np.floor(np.arange(out_h, dtype=F32) * sy).astype(np.int64), in_h - 1)
    xi = np.minimum(np.floor(np.arange(out_w, dtype=F32) * sx).astype(np.int64), in_w - 1)
    return x[..., yi[:, None], xi[None, :]]


def decouple_output(output, num_joints=17, num_stages=2, with_ae_loss=(True, False)):
    heat, tag = [], []
    for i in range(num_stages):
        heat.append(output[i][:, :num_joints])
        if with_ae_loss[i]:
            tag.append(output[i][:, num_joints:])
    return heat, tag


def parse_heatmaps(heat, tag, mask, num_stages=2):
    """-> (heatmap [N,K,H,W] aggregated + masked, tagging [N,K,H,W,T])."""
    if num_stages > 1:
        base = np.asarray(heat[-1], dtype=F32).copy()
        h, w = base.shape[-2:]
        for i in range(num_stages - 1):
            base = (base + resize_bilinear_legacy(heat[i], h, w)).astype(F32)
        base = (base / F32(num_stages)).astype(F32)
    else:
        base = np.asarray(heat[0], dtype=F32).copy()
        h, w = base.shape[-2:]
    tags = np.stack([resize_bilinear_legacy(t, h, w) for t in tag], axis=-1)
    m = resize_nearest(np.asarray(mask)[:, None].astype(F32), h, w) != 0
    base = np.where(m, base, F32(0)).astype(F32)
    return base, tags


def max_pool_same(x, k):
    lo = (k - 1) // 2
    hi = k - 1 - lo
    n, c, h, w = x.shape
    p = np.full((n, c, h + k - 1, w + k - 1), -np.inf, dtype=F32)
    p[:, :, lo:lo + h, lo:lo + w] = x
    out = np.full_like(x, -np.inf)
    for dy in range(k):
        for dx in range(k):
            out = np.maximum(out, p[:, :, dy:dy + h, dx:dx + w])
    del hi
    return out


def nms(heat, k):
    pooled = max_pool_same(heat, k)
    return (heat * (pooled == heat).astype(F32)).astype(F32)


def top_k(heat, tags, max_num, tag_per_joint=True):
    """-> val_k [N,K,M], tag_k [N,K,M,T], ind_k [N,K,M,2] (x, y), flat indices."""
    n, k, h, w = heat.shape
    flat = heat.reshape(n, k, -1)
    order = np.argsort(-flat, axis=2, kind="stable")[:, :, :max_num]
    val_k = np.take_along_axis(flat, order, axis=2)
    tflat = tags.reshape(n, tags.shape[1], h * w, -1)
    if not tag_per_joint:   # one tag plane for all joints (bottom_up_decoder.py:159-160)
        tflat = np.broadcast_to(tflat, (n, k) + tflat.shape[2:])
    tag_k = np.stack(
        [np.take_along_axis(tflat[..., t], order, axis=2) for t in range(tflat.shape[3])], axis=3)
    ind_k = np.stack((order % w, order // w), axis=3).astype(F32)
    return val_k.astype(F32), tag_k.astype(F32), ind_k, order


def shift_coordinate_quirk(ind_k, heat_raw, order):
    n, k, h, w = heat_raw.shape
    dx = np.zeros_like(heat_raw)
    dy = np.zeros_like(heat_raw)
    dx[:, :, :, 1:-1] = heat_raw[:, :, :, 2:] - heat_raw[:, :, :, :-2]
    dy[:, :, 1:-1, :] = heat_raw[:, :, 2:, :] - heat_raw[:, :, :-2, :]
    sx = np.sign(dx).reshape(n, k, -1)
    sy = np.sign(dy).reshape(n, k, -1)
    spatial = np.sort(order, axis=2)  # masked_select order: row-major, not rank order
    off_x = np.take_along_axis(sx, spatial, axis=2) * F32(0.25)
    off_y = np.take_along_axis(sy, spatial, axis=2) * F32(0.25)
    out = ind_k.copy()
    out[..., 0] += off_x.astype(F32)
    out[..., 1] += off_y.astype(F32)
    return out


def decode(model_output, mask, num_joints=17, num_stages=2, with_ae_loss=(True, False),
           use_nms=False, nms_kernel=5, max_num=30, shift_coordinate=False, tag_per_joint=True):
    """``BottomUpHeatMapAEDecoder.construct`` ->
    (val_k, tag_k, ind_k, heatmap_raw, tagging_heatmap)."""
    heat, tag = decouple_output(model_output, num_joints, num_stages, with_ae_loss)
    heatmap, tagging = parse_heatmaps(heat, tag, mask, num_stages)
    raw = heatmap.copy()
    if use_nms:
        heatmap = nms(heatmap, nms_kernel)
    val_k, tag_k, ind_k, order = top_k(heatmap, tagging, max_num, tag_per_joint)
    if shift_coordinate:
        ind_k = shift_coordinate_quirk(ind_k, raw, order)
    return val_k, tag_k, ind_k, raw, tagging
