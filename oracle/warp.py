"""Oracle (test infrastructure): OpenCV-exact affine crop warp.

The reference warps each crop with ``cv2.warpAffine(image, M, (w, h),
flags=cv2.INTER_LINEAR)`` (mindpose/data/transform/topdown_transform.py:217-222
and :248-253): forward matrix, BORDER_CONSTANT 0, uint8 HWC.  OpenCV is a
third-party dependency that is not under /root/reference (requirements.txt pins
``opencv-python>=4.2.0.34,<=4.5.4.60``; this image has 4.13.0), so its
published algorithm is restated here:

* the forward matrix is inverted in fp64 the way ``cv::warpAffine`` does;
* destination coordinates are mapped in 10-bit fixed point (AB_BITS = 10),
  ``adelta[x] = rint(Mi00 * x * 1024)``, ``X0 = rint((Mi01*y + Mi02) * 1024) + 16``
  and reduced to 1/32 pixel (INTER_BITS = 5);
* the four bilinear weights come from a 32x32 table of 15-bit integers
  (``rint(w * 32768)`` of the float32 products, INTER_REMAP_COEF_BITS = 15);
* taps outside the source read the constant border 0;
* ``dst = (sum(w * p) + 16384) >> 15`` (always inside [0, 255]).

PINNED: tests/golden/warp_*.npz hold ``cv2.warpAffine`` outputs generated in
the build container (oracle/gen_golden.py); tests/test_oracle_golden.py
requires this restatement to reproduce them bit for bit, and re-checks against
the installed cv2 when it is importable.
"""
import numpy as np

AB_BITS = 10
AB_SCALE = 1 << AB_BITS
INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
INTER_REMAP_COEF_BITS = 15
INTER_REMAP_COEF_SCALE = 1 << INTER_REMAP_COEF_BITS
ROUND_DELTA = AB_SCALE // INTER_TAB_SIZE // 2


def invert_affine(m):
    """fp64 inverse of a forward 2x3 matrix, op order of cv::warpAffine."""
    m = np.asarray(m, dtype=np.float64).reshape(2, 3)
    m00, m01, m02 = m[0]
    m10, m11, m12 = m[1]
    d = m00 * m11 - m01 * m10
    d = 1.0 / d if d != 0 else 0.0
    a11 = m11 * d
    a22 = m00 * d
    i00 = a11
    i01 = m01 * (-d)
    i10 = m10 * (-d)
    i11 = a22
    b1 = -i00 * m02 - i01 * m12
    b2 = -i10 * m02 - i11 * m12
    return np.array([[i00, i01, b1], [i10, i11, b2]], dtype=np.float64)


def _sat_int(v):
    """cv::saturate_cast<int>(double): round half to even, clamp to int32."""
    r = np.rint(v)
    return np.clip(r, -2147483648.0, 2147483647.0).astype(np.int64)


def bilinear_weight_table():
    """[32*32, 4] int32 weights (tl, tr, bl, br) indexed by fy*32+fx."""
    t = np.arange(INTER_TAB_SIZE, dtype=np.float32) * np.float32(1.0 / INTER_TAB_SIZE)
    one = np.float32(1.0)
    w1 = np.stack([one - t, t], axis=1)  # [32, 2] float32: (1-a, a)
    tab = np.zeros((INTER_TAB_SIZE, INTER_TAB_SIZE, 4), dtype=np.int32)
    for fy in range(INTER_TAB_SIZE):
        for fx in range(INTER_TAB_SIZE):
            w = np.array(
                [
                    w1[fy, 0] * w1[fx, 0],
                    w1[fy, 0] * w1[fx, 1],
                    w1[fy, 1] * w1[fx, 0],
                    w1[fy, 1] * w1[fx, 1],
                ],
                dtype=np.float32,
            )
            iw = np.clip(np.rint(w * np.float32(INTER_REMAP_COEF_SCALE)), -32768, 32767)
            iw = iw.astype(np.int32)
            # OpenCV redistributes a rounding residue onto the largest/smallest
            # weight so that the four always sum to 32768.
            diff = int(iw.sum()) - INTER_REMAP_COEF_SCALE
            if diff != 0:
                if diff < 0:
                    iw[int(np.argmax(iw))] -= diff
                else:
                    iw[int(np.argmin(iw))] -= diff
            tab[fy, fx] = iw
    return tab.reshape(-1, 4)


_WTAB = None


def warp_affine_u8(image, mat, dsize):
    """image: u8 [Hs, Ws, C]; mat: forward 2x3; dsize = (w, h). Returns u8 [h, w, C]."""
    global _WTAB
    if _WTAB is None:
        _WTAB = bilinear_weight_table()
    img = np.asarray(image)
    squeeze = img.ndim == 2
    if squeeze:
        img = img[:, :, None]
    hs, ws, ch = img.shape
    dw, dh = int(dsize[0]), int(dsize[1])
    mi = invert_affine(mat)

    xs = np.arange(dw, dtype=np.float64)
    ys = np.arange(dh, dtype=np.float64)
    adelta = _sat_int(mi[0, 0] * xs * AB_SCALE)
    bdelta = _sat_int(mi[1, 0] * xs * AB_SCALE)
    x0 = _sat_int((mi[0, 1] * ys + mi[0, 2]) * AB_SCALE) + ROUND_DELTA
    y0 = _sat_int((mi[1, 1] * ys + mi[1, 2]) * AB_SCALE) + ROUND_DELTA
    big_x = (x0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    big_y = (y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    sx = np.clip(big_x >> INTER_BITS, -32768, 32767)
    sy = np.clip(big_y >> INTER_BITS, -32768, 32767)
    fx = big_x & (INTER_TAB_SIZE - 1)
    fy = big_y & (INTER_TAB_SIZE - 1)
    w = _WTAB[(fy * INTER_TAB_SIZE + fx).astype(np.int64)]  # [dh, dw, 4]

    def tap(yy, xx):
        inside = (xx >= 0) & (xx < ws) & (yy >= 0) & (yy < hs)
        v = img[np.clip(yy, 0, hs - 1), np.clip(xx, 0, ws - 1)].astype(np.int64)
        return v * inside[:, :, None]

    acc = (
        tap(sy, sx) * w[:, :, 0:1]
        + tap(sy, sx + 1) * w[:, :, 1:2]
        + tap(sy + 1, sx) * w[:, :, 2:3]
        + tap(sy + 1, sx + 1) * w[:, :, 3:4]
    )
    out = (acc + (1 << (INTER_REMAP_COEF_BITS - 1))) >> INTER_REMAP_COEF_BITS
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out[:, :, 0] if squeeze else out


def normalize_chw(crop_u8, mean, std):
    """``vision.Normalize(mean, std)`` followed by ``vision.HWC2CHW()``
    (mindpose/data/data_factory.py:127-138; mean / std already multiplied by 255).

    PARITY UNPINNED: MindSpore's dataset ops cannot run here.  The arithmetic of record is
    ``(float32(pixel) - float32(mean[c])) / float32(std[c])`` in float32; MindSpore's
    implementation may instead multiply by a precomputed reciprocal (<= 1 ulp apart)."""
    mean = np.asarray(mean, np.float32)
    std = np.asarray(std, np.float32)
    out = (crop_u8.astype(np.float32) - mean[None, None, :]) / std[None, None, :]
    return np.ascontiguousarray(out.transpose(2, 0, 1))
