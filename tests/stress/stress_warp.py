"""Randomised parity stress of pc_warp_affine_u8 (3 channels: band kernel + list kernel) and
pc_warp_affine_u8_norm_chw against the oracle (the OpenCV fixed-point pipeline, pinned to
cv2.warpAffine goldens).  Development aid: python tests/stress/stress_warp.py [cases]

Random source sizes (widths that are and are not multiples of 4, so band and quad path both
run), images at odd byte offsets inside one buffer, destination widths that are multiples of
32, scales from 0.05 to 40, translations that push the crop over every edge, a third of the
crops rotated, a few mirrored."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mindpose_b200 import codec  # noqa: E402
from oracle import warp  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda", 0)
rng = np.random.RandomState(77)
bad = 0
for it in range(cases):
    n = int(rng.randint(1, 7))
    dw = int(rng.choice([32, 64, 96, 192, 288]))
    dh = int(rng.choice([1, 7, 16, 40, 64, 100, 256]))
    sizes = [(int(rng.randint(1, 90)), int(rng.choice([4, 8, 20, 52, 64, 100, 160, 161, 333])))
             for _ in range(n)]
    chunks, offs, imgs, pos = [], [], [], 0
    for hs, ws in sizes:
        pad = int(rng.randint(0, 7))
        chunks.append(rng.randint(0, 256, size=pad, dtype=np.uint8))
        pos += pad
        img = rng.randint(0, 256, size=(hs, ws, 3), dtype=np.uint8)
        offs.append(pos)
        imgs.append(img)
        chunks.append(img.reshape(-1))
        pos += img.size
    chunks.append(np.zeros(64, np.uint8))
    buf = np.concatenate(chunks)
    mats = np.zeros((n, 2, 3))
    for i, (hs, ws) in enumerate(sizes):
        s = float(np.exp(rng.uniform(np.log(0.05), np.log(40.0))))      # dst = s * src
        ang = np.deg2rad(rng.uniform(-60, 60)) if rng.rand() < 0.33 else 0.0
        sx = -s if rng.rand() < 0.05 else s
        mats[i] = [[sx * np.cos(ang), -s * np.sin(ang), rng.uniform(-0.6 * dw, 0.9 * dw)],
                   [sx * np.sin(ang), s * np.cos(ang), rng.uniform(-0.6 * dh, 0.9 * dh)]]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    inv = codec.invert_affine(t(mats))
    hw = torch.tensor(sizes, device=dev, dtype=torch.int32)
    out = codec.warp_affine(t(buf), torch.tensor(offs, device=dev), hw, inv, (dw, dh),
                            channels=3).cpu().numpy()
    mean, std = [123.7, 116.3, 103.5], [58.4, 57.1, 57.4]
    outf = codec.warp_affine_normalized(t(buf), torch.tensor(offs, device=dev), hw, inv, (dw, dh),
                                        mean, std).cpu().numpy()
    for i in range(n):
        want = warp.warp_affine_u8(imgs[i], mats[i], (dw, dh))
        ok = np.array_equal(out[i], want) and np.array_equal(outf[i], warp.normalize_chw(want, mean, std))
        if not ok:
            bad += 1
            print(f"MISMATCH case {it} crop {i}: src {sizes[i]} dst {(dh, dw)} mat {mats[i].tolist()} "
                  f"u8 diff {int((out[i] != want).sum())}")
print(f"{cases} cases, {bad} mismatching crops")
sys.exit(1 if bad else 0)
