"""Same bytes, half-height planes: 128 images of 64x128 + 128x256 maps against 64 images of
128x128 + 256x256 (one CTA per plane either way) -- what a 2-CTA split of a plane could gain."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mindpose_b200 as mp  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
dec = mp.create_decoder("bottomup_heatmap_ae", use_nms=True, nms_kernel=3, max_num=30)
dec.return_maps = False
for n, h in ((64, 256), (128, 128), (256, 64), (85, 192)):
    g = torch.Generator(device=dev).manual_seed(0)
    out0 = torch.rand(n, 34, h // 2, 128, device=dev, generator=g) * 0.02
    out1 = torch.rand(n, 17, h, 256, device=dev, generator=g) * 0.02
    ys = torch.randint(4, h - 4, (n, 17, 8), device=dev, generator=g)
    xs = torch.randint(8, 248, (n, 17, 8), device=dev, generator=g)
    for j in range(8):
        out1[torch.arange(n)[:, None], torch.arange(17)[None, :], ys[..., j], xs[..., j]] += 0.5
    mask = torch.ones(n, 2 * h, 512, dtype=torch.uint8, device=dev)
    for _ in range(3):
        dec([out0, out1], mask)
    ts = []
    for _ in range(15):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dec([out0, out1], mask)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"n {n} planes {n * 17} of {h} x 256: {ts[len(ts) // 2] * 1e3:.1f} us (min {ts[0] * 1e3:.1f})")
