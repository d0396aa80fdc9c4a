"""TEST INFRASTRUCTURE (oracle) -- not imported by the product.

CPU restatement of ``cv2.resize(image, (w, h), interpolation=cv2.INTER_LINEAR)`` for uint8
images, the arithmetic behind ``BottomUpRescale.transform``
(mindpose/data/transform/bottomup_transform.py:193-197), and of ``BottomUpRescale._get_new_size``
/ ``BottomUpPad.transform`` (:152-168, :610-640).

OpenCV is a third-party dependency of the reference (requirements.txt pins
opencv-python>=4.2.0.34,<=4.5.4.60); its bilinear resize is restated here from the published
algorithm (modules/imgproc/src/resize.cpp: the fixed-point HResizeLinear / VResizeLinear pair,
INTER_RESIZE_COEF_BITS = 11) and PINNED against cv2 itself (4.13.0 in this image) in
tests/test_oracle_resize.py, and against the unmodified reference classes through
tests/golden/bottomup_rescale_ref.npz (oracle/gen_golden_bottomup.py).

  * column dx: fx = float32((dx + 0.5) * (src_w / dst_w) - 0.5); sx = floor(fx); fx -= sx;
    sx < 0 -> (0, 0.0); sx >= src_w - 1 -> (src_w - 1, 0.0);
    weights int16 (round((1 - fx) * 2048), round(fx * 2048)), round half to even;
  * row dy: the same fy / sy WITHOUT zeroing fy; the two source rows are clamped to
    [0, src_h - 1] one by one;
  * horizontal pass in int32: r = S[sx] * a0 + S[sx + 1] * a1 (S[sx] * 2048 past the last column
    pair); vertical pass: ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2 >> 2;
  * an exact 2 x 2 reduction (src = 2 * dst in both directions) is the INTER_AREA fast path:
    (a + b + c + d + 2) >> 2.
"""
import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def _axis_table(src_n, dst_n, clamp_frac):
    """(index int64 [dst_n], w0 int32 [dst_n], w1 int32 [dst_n]) of one axis."""
    scale = 1.0 / (float(dst_n) / float(src_n))          # resize(): scale_x = 1. / inv_scale_x
    d = np.arange(dst_n, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_frac:
        lo = s < 0
        s[lo], f[lo] = 0, 0.0
        hi = s >= src_n - 1
        s[hi], f[hi] = src_n - 1, 0.0
    w0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)).astype(np.int32)
    w1 = np.rint(f * np.float32(COEF_SCALE)).astype(np.int32)
    return s, w0, w1


def resize_linear_u8(image, dst_wh):
    """image uint8 [H, W, C] -> uint8 [h, w, C] as cv2.resize(..., INTER_LINEAR) computes it."""
    image = np.ascontiguousarray(image)
    assert image.dtype == np.uint8 and image.ndim == 3
    sh, sw = image.shape[:2]
    dw, dh = int(dst_wh[0]), int(dst_wh[1])
    if sw == 2 * dw and sh == 2 * dh:                     # INTER_LINEAR -> INTER_AREA fast path
        s = image.astype(np.int32)
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    sx, a0, a1 = _axis_table(sw, dw, clamp_frac=True)
    sy, b0, b1 = _axis_table(sh, dh, clamp_frac=False)
    src = image.astype(np.int32)
    sx1 = np.minimum(sx + 1, sw - 1)                      # weight 0 where it was clamped
    rows = src[:, sx] * a0[None, :, None] + src[:, sx1] * a1[None, :, None]   # [H, w, C] int32
    y0 = np.clip(sy, 0, sh - 1)
    y1 = np.clip(sy + 1, 0, sh - 1)
    r0, r1 = rows[y0] >> 4, rows[y1] >> 4
    out = (((b0[:, None, None] * r0) >> 16) + ((b1[:, None, None] * r1) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def py_round(x):
    """Python's round() of a float: half to even, returns int."""
    return int(round(x))


def rescale_size(image_wh, max_wh):
    """BottomUpRescale._get_new_size (bottomup_transform.py:152-168)."""
    w, h = image_wh
    max_w, max_h = max_wh
    if w < h:
        max_w, max_h = max_h, max_w
    if w / h > max_w / max_h:
        return int(max_w), py_round(h * max_w / w)
    return py_round(w * max_h / h), int(max_h)


def rescale(image, max_wh, pixel_std=200.0):
    """BottomUpRescale.transform -> dict(image, center, scale, image_shape)."""
    h, w = image.shape[:2]
    tw, th = rescale_size((w, h), max_wh)
    return dict(image=resize_linear_u8(image, (tw, th)),
                center=np.array([py_round(w / 2), py_round(h / 2)]),
                scale=np.array([w / pixel_std, h / pixel_std]),
                image_shape=np.array([tw, th]))


def pad(image, max_wh):
    """BottomUpPad.transform -> dict(image, mask)."""
    h, w = image.shape[:2]
    tw, th = max_wh
    if w < h:
        th, tw = tw, th
    assert tw >= w and th >= h
    out = np.zeros((th, tw, image.shape[2]), image.dtype)
    out[:h, :w] = image
    mask = np.zeros((th, tw), np.uint8)
    mask[:h, :w] = 1
    return dict(image=out, mask=mask)
