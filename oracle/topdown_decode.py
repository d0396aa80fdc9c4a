"""Oracle (test infrastructure): top-down heatmap decoding and flip-test averaging.

float32 numpy restatement of two MindSpore graphs that cannot be executed here
(``mindspore`` is not installable in this image):

* ``TopDownHeatMapDecoder.construct`` and helpers
  (mindpose/models/decoders/top_down_decoder.py:72-215);
* the post-network half of ``_MultiRunNet.construct``
  (mindpose/engine/inferencer/topdown_inferencer.py:165-187).

PARITY UNPINNED.  The reference's tests for this path assert shapes only
(tests/models/decoders/test_top_down_decoder.py:8-47) and hold no numbers, so
this restatement is the oracle of record.  MindSpore 1.x semantics it assumes:

* ``ops.max(x, axis, keep_dims)`` returns ``(index, value)``; among equal
  maxima the lowest flat index is taken (== ``numpy.argmax``).
* ``_shift_coordinate`` (:118-141): the difference maps are zero on the border
  rows / columns because only the interior slice is assigned; ``sign(0) = 0``.
* ``ops.conv2d(pad_mode="same")`` with an 11x11 kernel zero-pads 5 pixels on
  each side and is a cross-correlation.  Its fp32 accumulation order is
  backend-defined; the order of record here is: for every kernel row, sum the
  products left to right; then sum the row sums top to bottom (each product and
  each partial sum rounded to float32, no fused multiply-add).
* ``ops.log`` is taken as the correctly rounded float32 logarithm
  (log in float64, rounded once).
* ``ops.pad`` (1.x signature) pads the LOG map with zeros (:178), so a peak on
  the border reads 0.0 for its outside neighbours.
* the flat fp32 index arithmetic of :181-183 is replaced by per-map integer
  indexing (identical for batches <= the reference's own ``batch_size: 128``;
  beyond that the reference loses integer precision in float32).
* ``MatrixInverse`` of the 2x2 ``H + 1e-7 * I`` is the adjugate formula in
  float32; the Einsum is ``inv @ d`` with products and one sum in float32.
* ``_transform_preds`` (:143-169): ``x * sx + c_x - s_w * 0.5`` evaluated left
  to right in float32, no contraction; ``sx = s_w / W`` (``W - 1`` for UDP).
* ``_flip_back`` / ``_shift_heatmap``: slice-assignment RHS is evaluated before
  the write, i.e. a true one-pixel shift with column 0 kept.
"""
import numpy as np

F32 = np.float32


def flip_index_from_pairs(flip_pairs):
    """topdown_inferencer.py:78-80 / topdown_transform.py:75-80."""
    idx = np.array(flip_pairs)[:, ::-1].flatten()
    return np.insert(idx, 0, 0)


def flip_average(heatmap, flipped, flip_index, shift_heatmap=False):
    """(heatmap + flip_back(flipped)) * 0.5  (topdown_inferencer.py:170-187)."""
    back = flipped[:, np.asarray(flip_index), ...]
    back = back[..., ::-1].copy()
    if shift_heatmap:
        shifted = back.copy()
        shifted[..., 1:] = back[..., :-1]
        back = shifted
    return ((heatmap.astype(F32) + back.astype(F32)) * F32(0.5)).astype(F32)


def max_preds(heatmap):
    """top_down_decoder.py:96-116 -> (coords f32 [N,K,2], maxvals f32 [N,K,1], idx)."""
    n, k, h, w = heatmap.shape
    flat = heatmap.reshape(n, k, -1)
    idx = np.argmax(flat, axis=2)
    maxvals = np.take_along_axis(flat, idx[..., None], axis=2)
    coords = np.zeros((n, k, 2), dtype=F32)
    coords[..., 0] = (idx % w).astype(F32)
    coords[..., 1] = np.floor(idx.astype(F32) / F32(w))
    return coords, maxvals.astype(F32), idx


def shift_coordinate(coords, heatmap, idx):
    """top_down_decoder.py:118-141."""
    n, k, h, w = heatmap.shape
    dx = np.zeros_like(heatmap)
    dy = np.zeros_like(heatmap)
    dx[:, :, :, 1:-1] = heatmap[:, :, :, 2:] - heatmap[:, :, :, :-2]
    dy[:, :, 1:-1, :] = heatmap[:, :, 2:, :] - heatmap[:, :, :-2, :]
    sx = np.sign(dx).reshape(n, k, -1)
    sy = np.sign(dy).reshape(n, k, -1)
    off_x = np.take_along_axis(sx, idx[..., None], axis=2)[..., 0] * F32(0.25)
    off_y = np.take_along_axis(sy, idx[..., None], axis=2)[..., 0] * F32(0.25)
    out = coords.copy()
    out[..., 0] += off_x.astype(F32)
    out[..., 1] += off_y.astype(F32)
    return out


def dark_gaussian_kernel(kernel_size=11):
    """top_down_decoder.py:207-215 -> float32 [ks, ks], normalised to sum 1."""
    sigma = 0.3 * ((kernel_size - 1) * 0.5 - 1) + 0.8
    xs = np.arange(-(kernel_size - 1) // 2, (kernel_size - 1) // 2 + 1, 1)
    ys = xs[:, None]
    kernel = np.exp(-(xs**2 + ys**2) / (2 * sigma**2))
    kernel = kernel / kernel.sum()
    return kernel.astype(F32)


def blur_same(heatmap, kernel):
    """Depthwise cross-correlation, zero 'same' padding, float32 accumulation in
    the order of record: row sums left->right, then rows top->bottom."""
    ks = kernel.shape[0]
    r = (ks - 1) // 2
    n, k, h, w = heatmap.shape
    padded = np.zeros((n, k, h + 2 * r, w + 2 * r), dtype=F32)
    padded[:, :, r : r + h, r : r + w] = heatmap
    total = None
    for ky in range(ks):
        row = None
        for kx in range(ks):
            prod = (kernel[ky, kx] * padded[:, :, ky : ky + h, kx : kx + w]).astype(F32)
            row = prod if row is None else (row + prod).astype(F32)
        total = row if total is None else (total + row).astype(F32)
    return total


def log_f32(x):
    return np.log(x.astype(np.float64)).astype(F32)


def dark_udp_refine(coords, heatmap, kernel):
    """top_down_decoder.py:171-205."""
    n, k, h, w = heatmap.shape
    blurred = blur_same(heatmap.astype(F32), kernel)
    blurred = np.clip(blurred, F32(0.001), F32(50))
    logmap = np.zeros((n, k, h + 2, w + 2), dtype=F32)
    logmap[:, :, 1:-1, 1:-1] = log_f32(blurred)

    px = coords[..., 0].astype(np.int64) + 1
    py = coords[..., 1].astype(np.int64) + 1
    nn = np.arange(n)[:, None]
    kk = np.arange(k)[None, :]

    def at(dy, dx):
        return logmap[nn, kk, py + dy, px + dx]

    i_ = at(0, 0)
    ix1 = at(0, 1)
    iy1 = at(1, 0)
    ix1y1 = at(1, 1)
    ix1_y1_ = at(-1, -1)
    ix1_ = at(0, -1)
    iy1_ = at(-1, 0)

    half = F32(0.5)
    two = F32(2)
    dx = half * (ix1 - ix1_)
    dy = half * (iy1 - iy1_)
    dxx = ix1 - two * i_ + ix1_
    dyy = iy1 - two * i_ + iy1_
    dxy = half * (ix1y1 - ix1 - iy1 + i_ + i_ - ix1_ - iy1_ + ix1_y1_)

    eps = F32(1e-7)
    a = dxx + eps
    b = dxy
    d = dyy + eps
    det = a * d - b * b
    i00 = d / det
    i01 = (-b) / det
    i11 = a / det
    off_x = i00 * dx + i01 * dy
    off_y = i01 * dx + i11 * dy
    out = coords.copy()
    out[..., 0] = coords[..., 0] - off_x
    out[..., 1] = coords[..., 1] - off_y
    return out.astype(F32)


def transform_preds(coords, center, scale, heatmap_hw, pixel_std=200.0, use_udp=False):
    """top_down_decoder.py:143-169."""
    h, w = heatmap_hw
    s = (scale.astype(F32) * F32(pixel_std)).astype(F32)
    if use_udp:
        sx = s[:, 0:1] / F32(w - 1.0)
        sy = s[:, 1:2] / F32(h - 1.0)
    else:
        sx = s[:, 0:1] / F32(w)
        sy = s[:, 1:2] / F32(h)
    out = np.ones_like(coords)
    out[:, :, 0] = coords[:, :, 0] * sx + center[:, 0:1] - s[:, 0:1] * F32(0.5)
    out[:, :, 1] = coords[:, :, 1] * sy + center[:, 1:2] - s[:, 1:2] * F32(0.5)
    return out.astype(F32)


def decode(
    heatmap,
    center,
    scale,
    score,
    pixel_std=200.0,
    to_original=True,
    shift_coordinate_flag=False,
    use_udp=False,
    dark_udp_refine_flag=False,
    kernel_size=11,
):
    """``TopDownHeatMapDecoder.construct`` -> (all_preds [N,K,3], all_boxes [N,6])."""
    if dark_udp_refine_flag and shift_coordinate_flag:
        raise ValueError(
            "`udp_refine` and `shift_coordinate` cannot be `true` in the same time."
        )
    heatmap = np.asarray(heatmap, dtype=F32)
    center = np.asarray(center, dtype=F32)
    scale = np.asarray(scale, dtype=F32)
    score = np.asarray(score, dtype=F32)
    n, k = heatmap.shape[:2]
    coords, maxvals, idx = max_preds(heatmap)
    if shift_coordinate_flag:
        coords = shift_coordinate(coords, heatmap, idx)
    elif dark_udp_refine_flag:
        coords = dark_udp_refine(coords, heatmap, dark_gaussian_kernel(kernel_size))
    if to_original:
        coords = transform_preds(
            coords, center, scale, heatmap.shape[2:], pixel_std, use_udp
        )
    all_preds = np.zeros((n, k, 3), dtype=F32)
    all_boxes = np.zeros((n, 6), dtype=F32)
    all_preds[:, :, 0:2] = coords[:, :, 0:2]
    all_preds[:, :, 2:3] = maxvals
    all_boxes[:, 0:2] = center[:, 0:2]
    all_boxes[:, 2:4] = scale[:, 0:2]
    sp = (scale * F32(pixel_std)).astype(F32)
    all_boxes[:, 4] = sp[:, 0] * sp[:, 1]
    all_boxes[:, 5] = score.reshape(n)
    return all_preds, all_boxes


def decode_with_flip(heatmap, flipped, flip_index, center, scale, score,
                     shift_heatmap=False, **decoder_kwargs):
    """Post-network half of the top-down ``_MultiRunNet.construct``."""
    final = flip_average(np.asarray(heatmap, F32), np.asarray(flipped, F32),
                         flip_index, shift_heatmap)
    return decode(final, center, scale, score, **decoder_kwargs)
