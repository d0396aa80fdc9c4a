"""The bilinear-resize oracle against cv2 itself, and the restated BottomUpRescale /
BottomUpPad / BottomUpResize host logic against the unmodified reference classes
(tests/golden/bottomup_rescale_ref.npz, made by oracle/gen_golden_bottomup.py)."""
import numpy as np
import pytest

import mindpose_b200 as mp
from mindpose_b200 import transforms
from oracle import gen_golden_bottomup as ggb
from oracle import resize as R

CFG = dict(image_size=[64, 64], max_image_size=ggb.RESCALE_MAX, heatmap_sizes=[[16, 16], [32, 32]],
           flip_pairs=[[1, 2]], pixel_std=200.0, tag_per_joint=True)


def test_resize_oracle_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(0)
    cases = [(480, 640, 512, 683), (640, 427, 767, 512), (100, 100, 50, 50), (101, 100, 50, 50),
             (37, 53, 91, 17), (5, 7, 64, 64), (2, 2, 9, 9), (1, 5, 3, 10), (9, 1, 4, 3),
             (64, 48, 128, 96), (600, 800, 300, 400)]
    cases += [tuple(int(v) for v in rng.randint(1, 200, 4)) for _ in range(60)]
    for sh, sw, dh, dw in cases:
        img = rng.randint(0, 256, (sh, sw, 3)).astype(np.uint8)
        want = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(R.resize_linear_u8(img, (dw, dh)), want), (sh, sw, dh, dw)


def test_rescale_and_pad_oracle_match_reference_golden(golden):
    g = golden("bottomup_rescale_ref.npz")
    for ci, (h, w) in enumerate(ggb.RESCALE_CASES):
        img = ggb.rescale_image(100 + ci, h, w)
        r = R.rescale(img, ggb.RESCALE_MAX)
        assert np.array_equal(r["image"], g[f"rescaled_{ci}"]), ci
        assert np.array_equal(r["center"], g[f"center_{ci}"])
        assert np.array_equal(r["scale"], g[f"scale_{ci}"])
        assert np.array_equal(r["image_shape"], g[f"shape_{ci}"])
        p = R.pad(r["image"], ggb.RESCALE_MAX)
        assert np.array_equal(p["image"], g[f"padded_{ci}"])
        assert np.array_equal(p["mask"], g[f"mask_{ci}"])


def test_host_side_of_the_transforms_matches_reference_golden(golden):
    """Target sizes, centres, scales (host arithmetic of the product) and BottomUpPad (a host
    copy) need no GPU."""
    g = golden("bottomup_rescale_ref.npz")
    pad = mp.create_transform("bottomup_pad", is_train=False, config=CFG)
    resize = transforms.BottomUpResize(is_train=False, config=CFG, size=64, base_length=32)
    for ci, (h, w) in enumerate(ggb.RESCALE_CASES):
        assert transforms._rescale_size((w, h), ggb.RESCALE_MAX) == tuple(g[f"shape_{ci}"])
        p = pad.transform(dict(image=g[f"rescaled_{ci}"]))
        assert np.array_equal(p["image"], g[f"padded_{ci}"])
        assert np.array_equal(p["mask"], g[f"mask_{ci}"])
        target, center, scale = resize._get_new_size((w, h), 200.0)
        assert tuple(target) == tuple(g[f"resized_shape_{ci}"])
        assert np.array_equal(center, g[f"resized_center_{ci}"])
        assert np.array_equal(scale, g[f"resized_scale_{ci}"])
    with pytest.raises(AssertionError):
        pad.transform(dict(image=np.zeros((70, 200, 3), np.uint8)))


def test_python_round_is_half_to_even_like_the_reference():
    assert transforms._rescale_size((101, 50), (104, 64)) == R.rescale_size((101, 50), (104, 64))
    for w, h in [(3, 2), (5, 2), (37, 33), (640, 427), (427, 640), (500, 375)]:
        assert transforms._rescale_size((w, h), (832, 512)) == R.rescale_size((w, h), (832, 512))


@pytest.mark.needs_reference
def test_sizes_and_images_against_live_reference():
    """In the build container: the unmodified BottomUpRescale / BottomUpResize / BottomUpPad for
    many image sizes (target sizes, centres, scales of the product's host arithmetic) and a few
    images through the oracle's arithmetic."""
    from oracle import ref_loader

    ns = ref_loader.load()
    cfg = dict(CFG, max_image_size=[832, 512])
    rescale = ns.bottomup.BottomUpRescale(is_train=False, config=cfg)
    resize = ns.bottomup.BottomUpResize(is_train=False, config=cfg, size=512, base_length=64)
    mine = transforms.BottomUpResize(is_train=False, config=cfg, size=512, base_length=64)
    rng = np.random.RandomState(1)
    sizes = [(640, 480), (480, 640), (500, 375), (333, 500), (1, 1), (832, 512), (1664, 1024)]
    sizes += [(int(rng.randint(1, 2000)), int(rng.randint(1, 2000))) for _ in range(300)]
    for w, h in sizes:
        assert transforms._rescale_size((w, h), cfg["max_image_size"]) == \
            rescale._get_new_size([w, h], cfg["max_image_size"]), (w, h)
        want = resize._get_new_size([w, h], 512, base_length=64, pixel_std=200.0)
        got = mine._get_new_size((w, h), 200.0)
        assert tuple(got[0]) == tuple(want[0]), (w, h)
        assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2]), (w, h)
    pad = ns.bottomup.BottomUpPad(is_train=False, config=cfg)
    for h, w in [(97, 131), (240, 427), (600, 400)]:
        img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        ref = rescale.transform(dict(image=img.copy()))
        got = R.rescale(img, cfg["max_image_size"])
        assert np.array_equal(got["image"], ref["image"]), (h, w)
        assert np.array_equal(got["center"], ref["center"]) and np.array_equal(got["scale"], ref["scale"])
        p_ref, p_got = pad.transform(dict(image=ref["image"])), R.pad(got["image"], cfg["max_image_size"])
        assert np.array_equal(p_got["image"], p_ref["image"]) and np.array_equal(p_got["mask"], p_ref["mask"])
