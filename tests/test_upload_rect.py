"""The rectangle PC_UPLOAD_ROI uploads for a crop (pc_crop_source_rect, host-only code of the
C-ABI library) must contain every source pixel the warp reads.  Checked on the CPU against
the oracle warp (which reads exactly OpenCV's taps): poisoning every pixel OUTSIDE the
rectangle must not change a single output byte.  The same rectangle logic runs inside
pc_topdown_affine_host; tests/test_topdown_gpu.py checks that path on the device."""
import ctypes

import numpy as np
import pytest

from mindpose_b200 import _lib
from oracle import affine, warp


def _rect(box, rot, hs, ws, image_size, use_udp, scale_padding=1.25):
    p = _lib.AffineHostParams(hs, ws, 3, image_size[0], image_size[1], 200.0, scale_padding,
                              int(use_udp), _lib.UPLOAD_ROI)
    box = np.asarray(box, np.float32)
    out = np.zeros(4, np.int32)
    _lib.call("pc_crop_source_rect", _lib.host_ptr(box), ctypes.c_float(rot), ctypes.byref(p),
              _lib.host_ptr(out))
    return [int(v) for v in out]


def _matrix(box, rot, image_size, use_udp):
    c, s = affine.box_to_center_scale(tuple(np.asarray(box, np.float32)), np.array(image_size))
    if use_udp:
        return affine.udp_matrix(c, s, rot, np.array(image_size))
    return affine.affine_matrix(c, s, rot, np.array(image_size))


@pytest.mark.parametrize("use_udp", [False, True])
@pytest.mark.parametrize("image_size", [(192, 256), (288, 384), (64, 48)])
def test_rectangle_contains_every_sampled_pixel(use_udp, image_size):
    rng = np.random.RandomState(hash((use_udp, image_size)) % (2 ** 31))
    hs, ws = 120, 160
    image = rng.randint(1, 256, size=(hs, ws, 3), dtype=np.uint8)
    saved = 0
    for case in range(60):
        bw, bh = rng.uniform(4, 150), rng.uniform(4, 110)
        box = [rng.uniform(-40, ws + 20 - bw), rng.uniform(-40, hs + 20 - bh), bw, bh]
        rot = 0.0 if case % 3 else float(rng.uniform(-180, 180))
        x0, y0, x1, y1 = _rect(box, rot, hs, ws, image_size, use_udp)
        assert 0 <= x0 and 0 <= y0 and x1 <= ws and y1 <= hs
        m = _matrix(box, rot, image_size, use_udp)
        want = warp.warp_affine_u8(image, m, image_size)
        poisoned = rng.randint(0, 256, size=image.shape, dtype=np.uint8)
        if x1 > x0 and y1 > y0:
            poisoned[y0:y1, x0:x1] = image[y0:y1, x0:x1]
            saved += image.size - (y1 - y0) * (x1 - x0) * 3
        got = warp.warp_affine_u8(poisoned, m, image_size)
        assert np.array_equal(got, want), (case, box, rot, (x0, y0, x1, y1))
    assert saved > 0   # the rectangle is not simply the whole image


@pytest.mark.parametrize("use_udp", [False, True])
def test_rectangle_for_extreme_boxes_and_right_angles(use_udp):
    """Tiny boxes (a few source pixels blown up to the whole crop), boxes much larger than
    the image, right-angle rotations, another scale_padding."""
    rng = np.random.RandomState(7 + use_udp)
    hs, ws = 72, 96
    image = rng.randint(1, 256, size=(hs, ws, 3), dtype=np.uint8)
    cases = [([40.3, 30.2, 1.0, 1.4], 0.0), ([10.0, 10.0, 2.5, 2.5], 37.0),
             ([-300.0, -200.0, 900.0, 700.0], 0.0), ([-300.0, -200.0, 900.0, 700.0], 45.0),
             ([20.0, 15.0, 50.0, 40.0], 90.0), ([20.0, 15.0, 50.0, 40.0], -90.0),
             ([20.0, 15.0, 50.0, 40.0], 180.0), ([20.0, 15.0, 50.0, 40.0], 270.0),
             ([0.0, 0.0, 96.0, 72.0], 0.0), ([95.0, 71.0, 3.0, 3.0], 0.0),
             ([-2.0, -2.0, 4.0, 4.0], 12.0), ([60.0, 5.0, 30.0, 66.0], -33.3)]
    for box, rot in cases:
        for pad in (1.25, 1.0):
            x0, y0, x1, y1 = _rect(box, rot, hs, ws, (48, 64), use_udp, scale_padding=pad)
            c, sc = affine.box_to_center_scale(tuple(np.asarray(box, np.float32)),
                                               np.array((48, 64)), scale_padding=pad)
            m = (affine.udp_matrix(c, sc, rot, np.array((48, 64))) if use_udp
                 else affine.affine_matrix(c, sc, rot, np.array((48, 64))))
            want = warp.warp_affine_u8(image, m, (48, 64))
            poisoned = rng.randint(0, 256, size=image.shape, dtype=np.uint8)
            if x1 > x0 and y1 > y0:
                poisoned[y0:y1, x0:x1] = image[y0:y1, x0:x1]
            assert np.array_equal(warp.warp_affine_u8(poisoned, m, (48, 64)), want), (box, rot, pad)


def test_rectangle_is_tight_for_an_axis_aligned_crop():
    # 100 x 133.33 box -> 125 x 166.7 source pixels around (150, 200)
    x0, y0, x1, y1 = _rect([100, 133.333, 100, 133.333], 0.0, 480, 640, (192, 256), False)
    assert 150 - 62.5 - 6 <= x0 <= 150 - 62.5 - 4 and 150 + 62.5 + 4 <= x1 <= 150 + 62.5 + 7
    assert 200 - 83.4 - 6 <= y0 <= 200 - 83.4 - 4 and 200 + 83.4 + 4 <= y1 <= 200 + 83.4 + 7


def test_rectangle_edge_cases():
    # a box far outside the image: empty rectangle (nothing to upload)
    x0, y0, x1, y1 = _rect([-900, -900, 50, 60], 0.0, 240, 320, (192, 256), False)
    assert x1 <= x0 or y1 <= y0
    # the whole image and more
    assert _rect([-50, -50, 500, 400], 0.0, 240, 320, (192, 256), False) == [0, 0, 320, 240]
    # not finite: refused (the front end falls back to the whole image)
    with pytest.raises(ValueError):
        _rect([float("nan"), 0, 10, 10], 0.0, 240, 320, (192, 256), False)
    with pytest.raises(ValueError):
        _rect([0, 0, float("inf"), 10], 0.0, 240, 320, (192, 256), False)
    # coordinates where the device's float32 centre / scale round by more than the padding allows
    with pytest.raises(ValueError):
        _rect([3e5, 0, 10, 10], 0.0, 240, 320, (192, 256), False)
    with pytest.raises(ValueError):
        _rect([0, 0, 2e5, 10], 0.0, 240, 320, (192, 256), False)
    # degenerate: the singular matrix's "inverse" samples pixel (0, 0), not the box
    with pytest.raises(ValueError):
        _rect([100, 100, 0, 0], 0.0, 240, 320, (192, 256), False)
