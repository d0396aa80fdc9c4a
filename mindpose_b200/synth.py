"""Seeded synthetic inputs of the BASELINE.json configs (numpy, host side).

Shared by tests/ and bench.py so that the CUDA path and the oracle always see
the same data.  Nothing here is a codec algorithm: these are input generators.
"""
from typing import Dict, Tuple

import numpy as np

COCO_FLIP_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]
COCO_JOINT_ORDER = [0, 1, 2, 3, 4, 5, 6, 11, 12, 7, 8, 9, 10, 13, 14, 15, 16]

TOPDOWN_CONFIG = dict(
    image_size=[192, 256],
    heatmap_size=[48, 64],
    pixel_std=200.0,
    scale_padding=1.25,
    flip_pairs=COCO_FLIP_PAIRS,
    upper_body_ids=list(range(11)),
)
TOPDOWN_CONFIG_384 = dict(TOPDOWN_CONFIG, image_size=[288, 384], heatmap_size=[72, 96])


def flip_index(pairs=COCO_FLIP_PAIRS) -> np.ndarray:
    return np.insert(np.array(pairs)[:, ::-1].flatten(), 0, 0)


def crop_geometry(n: int, seed: int = 0) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """center U(0,400), scale U(0.2,3), score U(0,1), the ranges the reference's
    decoder tests feed (tests/models/decoders/test_top_down_decoder.py:12-15)."""
    rng = np.random.RandomState(seed + 1000)
    center = rng.uniform(0, 400, size=(n, 2)).astype(np.float32)
    scale = rng.uniform(0.2, 3, size=(n, 2)).astype(np.float32)
    score = rng.random_sample(n).astype(np.float32)
    return center, scale, score


def noise_heatmaps(n: int, k: int, h: int, w: int, seed: int = 0) -> np.ndarray:
    return np.random.RandomState(seed).random_sample((n, k, h, w)).astype(np.float32)


def blob_heatmaps(n: int, k: int, h: int, w: int, seed: int = 0, sigma: float = 2.0,
                  noise: float = 0.02) -> Tuple[np.ndarray, np.ndarray]:
    """One Gaussian per joint map: sub-pixel centre in U(3, W-4) x U(3, H-4),
    amplitude U(0.3, 1), plus U(0, noise).  Returns (maps, centres [n,k,2])."""
    rng = np.random.RandomState(seed)
    cx = rng.uniform(3, w - 4, size=(n, k)).astype(np.float32)
    cy = rng.uniform(3, h - 4, size=(n, k)).astype(np.float32)
    amp = rng.uniform(0.3, 1.0, size=(n, k)).astype(np.float32)
    xs = np.arange(w, dtype=np.float32)[None, None, None, :]
    ys = np.arange(h, dtype=np.float32)[None, None, :, None]
    d2 = (xs - cx[..., None, None]) ** 2 + (ys - cy[..., None, None]) ** 2
    maps = amp[..., None, None] * np.exp(-d2 / np.float32(2 * sigma * sigma))
    maps = maps + rng.uniform(0, noise, size=maps.shape)
    return maps.astype(np.float32), np.stack([cx, cy], axis=-1)


def flipped_pair(maps: np.ndarray, seed: int = 0, noise: float = 0.02,
                 fidx: np.ndarray = None) -> np.ndarray:
    """A plausible 'network output on the mirrored image': mirror + joint swap of
    ``maps`` plus independent noise."""
    fidx = flip_index() if fidx is None else fidx
    rng = np.random.RandomState(seed + 77)
    inv = np.argsort(fidx)
    mirrored = maps[:, inv][..., ::-1]
    return (mirrored + rng.uniform(0, noise, size=maps.shape)).astype(np.float32)


def keypoints(n: int, k: int, image_size, seed: int = 0) -> np.ndarray:
    """x in U(-8, w+8), y in U(-8, h+8): some windows clip, some fall outside;
    visibility Bernoulli(0.8)."""
    rng = np.random.RandomState(seed)
    w, h = image_size
    kp = np.zeros((n, k, 3), dtype=np.float32)
    kp[..., 0] = rng.uniform(-8, w + 8, size=(n, k))
    kp[..., 1] = rng.uniform(-8, h + 8, size=(n, k))
    kp[..., 2] = (rng.random_sample((n, k)) < 0.8).astype(np.float32)
    return kp


def source_images_and_boxes(n: int, hs: int = 480, ws: int = 640, seed: int = 0
                            ) -> Tuple[np.ndarray, np.ndarray]:
    """n noise images u8 [hs, ws, 3] and one (x, y, w, h) box inside each."""
    rng = np.random.RandomState(seed)
    images = rng.randint(0, 256, size=(n, hs, ws, 3), dtype=np.uint8)
    bw = rng.uniform(40, min(400, ws - 1), size=n)
    bh = rng.uniform(60, min(440, hs - 1), size=n)
    bx = rng.uniform(0, 1, size=n) * (ws - bw)
    by = rng.uniform(0, 1, size=n) * (hs - bh)
    boxes = np.stack([bx, by, bw, bh], axis=1).astype(np.float32)
    return images, boxes


def bottomup_outputs(n: int, k: int = 17, h0: int = 128, w0: int = 128, mask_hw=(512, 512),
                     seed: int = 0, max_people: int = 12) -> Dict[str, np.ndarray]:
    """HigherHRNet-style outputs: out0 [n,2k,h0,w0] (heat | tag), out1 [n,k,2h0,2w0],
    mask u8 [n,Hm,Wm] (1 = valid, with one zero rectangle)."""
    rng = np.random.RandomState(seed)
    h1, w1 = 2 * h0, 2 * w0
    out0 = np.zeros((n, 2 * k, h0, w0), dtype=np.float32)
    out1 = np.zeros((n, k, h1, w1), dtype=np.float32)
    ys0, xs0 = np.mgrid[0:h0, 0:w0].astype(np.float32)
    ys1, xs1 = np.mgrid[0:h1, 0:w1].astype(np.float32)
    for i in range(n):
        out0[i, :k] = rng.uniform(0, 0.02, size=(k, h0, w0))
        out0[i, k:] = rng.uniform(-1, 1, size=(k, h0, w0))
        out1[i] = rng.uniform(0, 0.02, size=(k, h1, w1))
        people = rng.randint(1, max_people + 1)
        for pid in range(people):
            base = rng.uniform(20, [w1 - 20, h1 - 20])
            for j in range(k):
                if rng.random_sample() < 0.15:
                    continue
                cx, cy = base + rng.uniform(-18, 18, size=2)
                amp = rng.uniform(0.4, 1.0)
                out1[i, j] += amp * np.exp(-((xs1 - cx) ** 2 + (ys1 - cy) ** 2) / 8.0)
                out0[i, j] += amp * np.exp(-((xs0 - cx / 2) ** 2 + (ys0 - cy / 2) ** 2) / 8.0)
                near = (xs0 - cx / 2) ** 2 + (ys0 - cy / 2) ** 2 < 16.0
                tagv = pid * 3.0 + rng.normal(0, 0.1, size=(h0, w0))
                out0[i, k + j][near] = tagv[near]
    mask = np.ones((n, mask_hw[0], mask_hw[1]), dtype=np.uint8)
    for i in range(n):
        y0 = rng.randint(0, mask_hw[0] - 64)
        x0 = rng.randint(0, mask_hw[1] - 64)
        mask[i, y0:y0 + rng.randint(16, 64), x0:x0 + rng.randint(16, 64)] = 0
    return dict(out0=out0.astype(np.float32), out1=out1.astype(np.float32), mask=mask)
