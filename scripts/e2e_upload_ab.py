#!/usr/bin/env python
"""A/B of the host-buffer crop warp's upload modes (PC_UPLOAD_FULL / ROI / ROI_KERNEL) on the
bench.py workload: pinned host images u8 [n,480,640,3] -> host crops u8 [n,256,192,3].
Prints one JSON line per mode (crops/s, bytes over PCIe, GB/s achieved host->device).

    python scripts/e2e_upload_ab.py [--crops 2048] [--reps 3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--crops", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch

    from mindpose_b200 import codec, synth

    n, hs, ws = args.crops, 480, 640
    cfg = synth.TOPDOWN_CONFIG
    g = torch.Generator().manual_seed(0)
    base = torch.randint(0, 256, (128, hs, ws, 3), dtype=torch.uint8, generator=g)
    images = base.repeat((n + 127) // 128, 1, 1, 1)[:n].contiguous().pin_memory()
    rng = np.random.RandomState(0)      # the box distribution of bench.py
    bw, bh = rng.uniform(40, 400, n), rng.uniform(60, 440, n)
    boxes = np.stack([rng.uniform(0, 1, n) * (ws - bw), rng.uniform(0, 1, n) * (hs - bh), bw, bh],
                     axis=1).astype(np.float32)
    crops = torch.empty(n, 256, 192, 3, dtype=torch.uint8).pin_memory()
    ctx = codec.HostContext(0, scratch_bytes=2 << 30)
    ref = None
    variants = ["full", "roi", "roi_kernel"]
    for mode in variants * 3:
        kw = dict(out=crops.numpy(), upload=mode)
        ctx.topdown_affine(images.numpy(), boxes, cfg["image_size"], **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            ctx.topdown_affine(images.numpy(), boxes, cfg["image_size"], **kw)
        dt = (time.perf_counter() - t0) / args.reps
        h2d, d2h = ctx.last_transfer_bytes()
        same = None
        if ref is None:
            ref = crops.numpy().copy()
        else:
            same = bool(np.array_equal(ref, crops.numpy()))
        print(json.dumps({"upload": mode, "crops": n, "ms": dt * 1e3, "crops_per_s": n / dt,
                          "h2d_bytes": h2d, "d2h_bytes": d2h, "h2d_gbs": h2d / dt / 1e9,
                          "same_crops_as_full": same}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
