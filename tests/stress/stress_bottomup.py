"""Randomised parity stress of pc_bottomup_decode (pair kernel shapes) against the oracle.
Development aid: python tests/stress/stress_bottomup.py [cases]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mindpose_b200 as mp  # noqa: E402
from mindpose_b200 import bottomup, synth  # noqa: E402
from oracle import bottomup_decode as bd  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda", 0)
rng = np.random.RandomState(2024)
bad = 0
for it in range(cases):
    w0 = int(rng.choice([68, 72, 80, 96, 100, 112, 128]))       # W = 2 w0 in (128, 256]
    h0 = int(rng.choice([8, 17, 30, 48, 64, 96, 128, 150]))
    n = int(rng.randint(1, 3))
    k = int(rng.choice([1, 5, 17]))
    mh = 2 * h0 * int(rng.choice([1, 2])) + int(rng.choice([0, 0, 6]))
    mw = 4 * w0
    kind = rng.choice(["people", "noise", "sparse", "plateau", "negative"])
    if kind == "people" and (mh < 70 or mw < 70):
        kind = "noise"          # the generator wants room for its mask rectangle
    if kind == "people":
        d = synth.bottomup_outputs(n, k, h0, w0, mask_hw=(mh, mw), seed=int(rng.randint(1 << 30)),
                                   max_people=int(rng.randint(1, 20)))
        out0, out1, mask = d["out0"], d["out1"], d["mask"]
    else:
        out0 = rng.uniform(-0.02, 0.02, (n, 2 * k, h0, w0)).astype(np.float32)
        out1 = rng.uniform(-0.02, 0.02, (n, k, 2 * h0, 2 * w0)).astype(np.float32)
        mask = np.ones((n, mh, mw), np.uint8)
        if kind == "sparse":
            out0[:, :k] = 0
            out1[:] = 0
            for _ in range(int(rng.randint(0, 60))):
                out1[rng.randint(n), rng.randint(k), rng.randint(2 * h0), rng.randint(2 * w0)] = \
                    rng.uniform(0.1, 1)
        elif kind == "plateau":
            out1[:, ::2] = np.round(out1[:, ::2] * 200) / 200      # heavy ties
            out0[:, :k:2] = 0.25
        elif kind == "negative":
            out0[:, :k] -= 0.5
            out1 -= 0.5
    for _ in range(int(rng.randint(0, 5))):
        y0, x0 = rng.randint(0, mh - 2), rng.randint(0, mw - 2)
        mask[rng.randint(n), y0:y0 + rng.randint(1, 60), x0:x0 + rng.randint(1, 200)] = 0
    use_nms = bool(rng.rand() < 0.8)
    m = int(rng.choice([30, 30, 12, 32, 1]))
    want = bd.decode([out0, out1], mask, num_joints=k, use_nms=use_nms, nms_kernel=3, max_num=m)
    dec = mp.create_decoder("bottomup_heatmap_ae", num_joints=k, use_nms=use_nms, nms_kernel=3,
                            max_num=m)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    bottomup.decode_stats(reset=True)
    got = dec([t(out0), t(out1)], t(mask))
    ok = all(np.array_equal(g.cpu().numpy(), w) for g, w in zip(got, want))
    exact = bottomup.decode_stats()
    print(f"case {it:3d} {kind:9s} n={n} k={k} H={2*h0} W={2*w0} mask={mh}x{mw} nms={use_nms} M={m} "
          f"exact-pass planes {exact}/{n*k}: {'ok' if ok else 'MISMATCH'}", flush=True)
    bad += 0 if ok else 1
print("mismatches:", bad)
sys.exit(1 if bad else 0)
