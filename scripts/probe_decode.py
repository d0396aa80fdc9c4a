"""Run one decode config in isolation and report the CUDA error text (debug aid)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import mindpose_b200 as mp
from mindpose_b200 import synth

n, mode, flip, data, iters = int(sys.argv[1]), sys.argv[2], sys.argv[3] == "flip", sys.argv[4], int(sys.argv[5])
dev = torch.device("cuda", 0)
k, h, w = 17, 64, 48
if data == "noise":
    hm = torch.rand(n, k, h, w, device=dev); fl = torch.rand(n, k, h, w, device=dev)
else:
    hm = torch.rand(n, k, h, w, device=dev) * 0.02; fl = torch.rand(n, k, h, w, device=dev) * 0.02
    hm[:, :, 30, 20] = 0.9; fl[:, :, 30, 27] = 0.9
center = torch.rand(n, 2, device=dev) * 400
scale = torch.rand(n, 2, device=dev) * 2.8 + 0.2
score = torch.rand(n, device=dev)
kw = dict(plain=dict(to_original=False), orig=dict(), shift=dict(shift_coordinate=True), dark=dict(dark_udp_refine=True))[mode]
dec = mp.create_decoder("topdown_heatmap", **kw)
try:
    for it in range(iters):
        if flip:
            p, b = dec.decode_flip_pair(hm, fl, synth.flip_index(), center, scale, score)
        else:
            p, b = dec(hm, center, scale, score)
        torch.cuda.synchronize()
    print("OK", sys.argv[1:], flush=True)
except Exception as e:
    print("FAIL", sys.argv[1:], type(e).__name__, str(e).splitlines()[0][:200], flush=True)
