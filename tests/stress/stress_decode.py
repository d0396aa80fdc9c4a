"""Stress the top-down decode kernel: many launches, every mode, sync + progress after each."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import mindpose_b200 as mp
from mindpose_b200 import synth

dev = torch.device("cuda", 0)
n, k = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 17
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
for (h, w) in ((64, 48), (96, 72)):
    hm = torch.rand(n, k, h, w, device=dev)
    fl = torch.rand(n, k, h, w, device=dev)
    center = torch.rand(n, 2, device=dev) * 400
    scale = torch.rand(n, 2, device=dev) * 2.8 + 0.2
    score = torch.rand(n, device=dev)
    for kwargs in (dict(to_original=False), dict(shift_coordinate=True), dict(dark_udp_refine=True)):
        dec = mp.create_decoder("topdown_heatmap", **kwargs)
        for flip in (True, False):
            t0 = time.time()
            for it in range(iters):
                if flip:
                    p, b = dec.decode_flip_pair(hm, fl, synth.flip_index(), center, scale, score)
                else:
                    p, b = dec(hm, center, scale, score)
                torch.cuda.synchronize()
            print(h, w, kwargs, "flip" if flip else "noflip", "ok", f"{(time.time()-t0)/iters*1e3:.3f} ms/launch", flush=True)
print("stress done", flush=True)
