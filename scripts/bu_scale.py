"""Bottom-up decode time against the number of planes in flight (n images x 17 joints, one CTA
per plane, 4 CTAs per SM): tells whether a CTA's run time depends on how many others share the
SM (throughput bound) or not (latency bound).  python scripts/bu_scale.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mindpose_b200 as mp  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
nmax = 256
out0 = torch.rand(nmax, 34, 128, 128, device=dev, generator=g) * 0.02
out1 = torch.rand(nmax, 17, 256, 256, device=dev, generator=g) * 0.02
ys = torch.randint(8, 248, (nmax, 17, 8), device=dev, generator=g)
xs = torch.randint(8, 248, (nmax, 17, 8), device=dev, generator=g)
for j in range(8):
    out1[torch.arange(nmax)[:, None], torch.arange(17)[None, :], ys[..., j], xs[..., j]] += 0.5
mask = torch.ones(nmax, 512, 512, dtype=torch.uint8, device=dev)
dec = mp.create_decoder("bottomup_heatmap_ae", use_nms=True, nms_kernel=3, max_num=30)
dec.return_maps = False
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n in (4, 8, 17, 26, 34, 35, 52, 64, 69, 70, 104, 128, 256):
    a, b, m = out0[:n], out1[:n], mask[:n]
    for _ in range(3):
        dec([a, b], m)
    ts = []
    for _ in range(15):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dec([a, b], m)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    t = ts[len(ts) // 2]
    planes = n * 17
    print(f"n {n:4d} planes {planes:5d} ({planes / 592:.2f} waves of 592)  {t * 1e3:7.1f} us  "
          f"{t * 1e3 / max(1, -(-planes // 592)):.1f} us per wave  {n * 6.9e6 / (t * 1e-3) / 1e9:.0f} GB/s")
