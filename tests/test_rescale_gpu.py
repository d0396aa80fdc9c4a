"""Bottom-up evaluation preprocessing on the device (pc_rescale_pad_u8, BottomUpRescale /
BottomUpPad / BottomUpResize) against the unmodified reference's outputs
(tests/golden/bottomup_rescale_ref.npz) and the cv2-pinned oracle."""
import hashlib

import numpy as np
import pytest
import torch

import mindpose_b200 as mp
from mindpose_b200 import codec
from oracle import gen_golden_bottomup as ggb
from oracle import resize as R

pytestmark = pytest.mark.gpu

CFG = dict(image_size=[64, 64], max_image_size=ggb.RESCALE_MAX, heatmap_sizes=[[16, 16], [32, 32]],
           flip_pairs=[[1, 2]], pixel_std=200.0, tag_per_joint=True)


def test_rescale_pad_batch_matches_reference_golden(cuda_device, golden):
    g = golden("bottomup_rescale_ref.npz")
    t = mp.create_transform("bottomup_rescale", is_train=False, config=CFG)
    land = [ci for ci, (h, w) in enumerate(ggb.RESCALE_CASES) if w >= h]
    port = [ci for ci, (h, w) in enumerate(ggb.RESCALE_CASES) if w < h]
    for ids, canvas in ((land, tuple(ggb.RESCALE_MAX)), (port, tuple(ggb.RESCALE_MAX[::-1]))):
        imgs = [ggb.rescale_image(100 + ci, *ggb.RESCALE_CASES[ci]) for ci in ids]
        out, mask, meta = t.rescale_pad_batch(imgs, canvas_wh=canvas)
        out, mask = out.cpu().numpy(), mask.cpu().numpy()
        for j, ci in enumerate(ids):
            assert np.array_equal(out[j], g[f"padded_{ci}"]), ci
            assert np.array_equal(mask[j], g[f"mask_{ci}"]), ci
            assert np.array_equal(meta["center"][j], g[f"center_{ci}"])
            assert np.array_equal(meta["scale"][j], g[f"scale_{ci}"])
            assert np.array_equal(meta["image_shape"][j], g[f"shape_{ci}"])


def test_single_sample_transforms_match_reference_golden(cuda_device, golden):
    g = golden("bottomup_rescale_ref.npz")
    rescale = mp.create_transform("bottomup_rescale", is_train=False, config=CFG)
    pad = mp.create_transform("bottomup_pad", is_train=False, config=CFG)
    resize = mp.create_transform("bottomup_resize", is_train=False, config=CFG, size=64,
                                 base_length=32)
    for ci, (h, w) in enumerate(ggb.RESCALE_CASES):
        img = ggb.rescale_image(100 + ci, h, w)
        r = rescale.transform(dict(image=img.copy()))
        assert np.array_equal(r["image"], g[f"rescaled_{ci}"]), ci
        assert np.array_equal(r["center"], g[f"center_{ci}"])
        assert np.array_equal(r["scale"], g[f"scale_{ci}"])
        assert tuple(r["image_shape"]) == tuple(g[f"shape_{ci}"])
        p = pad.transform(dict(image=r["image"]))
        assert np.array_equal(p["image"], g[f"padded_{ci}"])
        assert np.array_equal(p["mask"], g[f"mask_{ci}"])
        # the column-tuple calling convention of the dataset pipeline
        z = resize.transform(dict(image=img.copy()))
        assert np.array_equal(z["image"], g[f"resized_{ci}"]), ci
        assert np.array_equal(z["mask"], g[f"resized_mask_{ci}"])
        assert np.array_equal(z["center"], g[f"resized_center_{ci}"])
        assert np.array_equal(z["scale"], g[f"resized_scale_{ci}"])
        assert tuple(z["image_shape"]) == tuple(g[f"resized_shape_{ci}"])


def test_shipped_recipe_sizes_match_reference_digests(cuda_device, golden):
    g = golden("bottomup_rescale_ref.npz")
    big = dict(CFG, max_image_size=[832, 512])
    t = mp.create_transform("bottomup_rescale", is_train=False, config=big)
    for ci, (h, w) in enumerate([(480, 640), (640, 427), (375, 500)]):
        canvas = (832, 512) if w >= h else (512, 832)
        out, mask, _ = t.rescale_pad_batch([ggb.rescale_image(200 + ci, h, w)], canvas_wh=canvas)
        digest = hashlib.sha256(out[0].cpu().numpy().tobytes()
                                + mask[0].cpu().numpy().tobytes()).hexdigest()
        assert digest == str(g["big_digests"][ci]), ci


def test_rescale_pad_random_sizes_match_oracle(cuda_device):
    """Ragged batch: up- and down-scaling, the exact 2 x 2 reduction, one-pixel sources, a
    canvas width that is not a multiple of 4, targets that fill the canvas exactly."""
    rng = np.random.RandomState(3)
    dev = cuda_device
    for cw, ch in ((104, 64), (103, 61), (257, 130)):
        sizes, targets = [], []
        for i in range(24):
            sh, sw = int(rng.randint(1, 150)), int(rng.randint(1, 150))
            tw, th = int(rng.randint(1, cw + 1)), int(rng.randint(1, ch + 1))
            if i == 0:
                sh, sw, tw, th = 2 * (ch // 2), 2 * (cw // 2), cw // 2, ch // 2    # area path
            if i == 1:
                tw, th = cw, ch
            if i == 2:
                sh, sw = 1, 1
            sizes.append((sh, sw))
            targets.append((tw, th))
        imgs = [rng.randint(0, 256, (sh, sw, 3)).astype(np.uint8) for sh, sw in sizes]
        offs, total = [], 0
        for sh, sw in sizes:
            offs.append(total)
            total += (sh * sw * 3 + 15) // 16 * 16
        blob = np.zeros(total, np.uint8)
        for im, off in zip(imgs, offs):
            blob[off:off + im.size] = im.reshape(-1)
        out, mask = codec.rescale_pad(torch.from_numpy(blob).to(dev),
                                      torch.tensor(offs, dtype=torch.int64),
                                      torch.tensor(sizes, dtype=torch.int32),
                                      torch.tensor(targets, dtype=torch.int32), (cw, ch))
        out, mask = out.cpu().numpy(), mask.cpu().numpy()
        for i, (im, (tw, th)) in enumerate(zip(imgs, targets)):
            want = np.zeros((ch, cw, 3), np.uint8)
            want[:th, :tw] = R.resize_linear_u8(im, (tw, th))
            wm = np.zeros((ch, cw), np.uint8)
            wm[:th, :tw] = 1
            assert np.array_equal(out[i], want), (cw, ch, i, sizes[i], targets[i])
            assert np.array_equal(mask[i], wm), (cw, ch, i)


def test_rescale_pad_argument_errors(cuda_device):
    t = mp.create_transform("bottomup_rescale", is_train=False, config=CFG)
    with pytest.raises(ValueError, match="does not fit"):
        t.rescale_pad_batch([np.zeros((80, 60, 3), np.uint8)])       # portrait, landscape canvas
    with pytest.raises(ValueError, match="3 channels"):
        t.rescale_pad_batch([np.zeros((8, 8, 1), np.uint8)])
    out, mask, meta = t.rescale_pad_batch([], canvas_wh=(8, 8))
    assert out.shape == (0, 8, 8, 3) and mask.shape == (0, 8, 8)


@pytest.mark.parametrize("stats", ["imagenet", "other"])
def test_rescale_pad_normalized_chw(cuda_device, golden, stats):
    """Normalize + HWC2CHW fused in: equal to (float32(canvas) - mean) / std on the bit-exact
    uint8 canvas of the reference, padding included (Normalize sees the padded zeros too)."""
    g = golden("bottomup_rescale_ref.npz")
    t = mp.create_transform("bottomup_rescale", is_train=False, config=CFG)
    if stats == "imagenet":
        mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    else:
        mean, std = [0.31, 0.5, 0.123], [0.3, 0.11, 0.27]     # (the division path)
    ids = [ci for ci, (h, w) in enumerate(ggb.RESCALE_CASES) if w >= h]
    imgs = [ggb.rescale_image(100 + ci, *ggb.RESCALE_CASES[ci]) for ci in ids]
    for canvas in (tuple(ggb.RESCALE_MAX), (105, 67)):
        out, mask, _ = t.rescale_pad_batch(imgs, canvas_wh=canvas, normalize_mean=mean,
                                           normalize_std=std)
        out, mask = out.cpu().numpy(), mask.cpu().numpy()
        m32 = (np.array(mean) * 255.0).astype(np.float32)
        s32 = (np.array(std) * 255.0).astype(np.float32)
        for j, ci in enumerate(ids):
            ref = g[f"padded_{ci}"]
            u8 = np.zeros((canvas[1], canvas[0], 3), np.uint8)
            u8[:ref.shape[0], :ref.shape[1]] = ref
            want = ((u8.astype(np.float32) - m32) / s32).transpose(2, 0, 1)
            assert out[j].dtype == np.float32 and np.array_equal(out[j], want), (stats, canvas, ci)
            wm = np.zeros((canvas[1], canvas[0]), np.uint8)
            wm[:ref.shape[0], :ref.shape[1]] = g[f"mask_{ci}"]
            assert np.array_equal(mask[j], wm)


def test_rescale_pad_camera_sizes_match_oracle(cuda_device):
    """Camera-sized sources into the recipe's canvas: 2.3 x and 3.6 x reductions (no source row
    is shared between output rows), the exact 2 x 2 reduction, a row pitch that is not a
    multiple of 4 bytes, a portrait image in the turned canvas."""
    big = dict(CFG, max_image_size=[832, 512])
    t = mp.create_transform("bottomup_rescale", is_train=False, config=big)
    rng = np.random.RandomState(9)
    for (h, w), canvas in (((1080, 1920), (832, 512)), ((1024, 1664), (832, 512)),
                           ((1333, 2001), (832, 512)), ((3000, 2000), (512, 832))):
        img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        out, mask, meta = t.rescale_pad_batch([img], canvas_wh=canvas)
        tw, th = R.rescale_size((w, h), (832, 512))
        assert tuple(meta["image_shape"][0]) == (tw, th)
        want = np.zeros((canvas[1], canvas[0], 3), np.uint8)
        want[:th, :tw] = R.resize_linear_u8(img, (tw, th))
        assert np.array_equal(out[0].cpu().numpy(), want), (h, w)
        assert int(mask[0].sum()) == tw * th
