"""Randomised parity stress of pc_oks_nms against the oracle (hard + soft, ties, iou vis_thr).
Development aid: python tests/stress/stress_nms.py [cases]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mindpose_b200 import nms as dnms  # noqa: E402
from oracle import gen_golden_nms as ggn  # noqa: E402
from oracle import nms as onms  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda", 0)
rng = np.random.RandomState(7)
bad = 0
for it in range(cases):
    images = int(rng.randint(1, 6))
    counts = [int(rng.choice([0, 1, 2, 5, 13, 40, 90])) for _ in range(images)]
    thr = float(rng.choice([0.3, 0.5, 0.8, 0.9, 0.95]))
    soft = bool(rng.rand() < 0.5)
    vthr = None if rng.rand() < 0.7 else float(rng.uniform(0.1, 0.6))
    rvis = float(rng.choice([0.0, 0.2, 0.5]))
    ks, ars, scs = [], [], []
    for i, c in enumerate(counts):
        k, a, s = ggn.nms_people(int(rng.randint(1 << 30)), max(c, 1))
        if rng.rand() < 0.3:
            s = np.round(s * 8).astype(np.float32) / 8          # score ties
        ks.append(k[:c]), ars.append(a[:c]), scs.append(s[:c])
    if sum(counts) == 0:
        continue
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    score = t(np.concatenate(scs).astype(np.float32))
    keep, num = dnms.rescore_and_nms(t(np.concatenate(ks)), t(np.concatenate(ars)), score, t(off),
                                     max(counts), oks_thr=thr, rescore_vis_thr=rvis, soft=soft,
                                     max_dets=20, iou_vis_thr=vthr)
    keep, num, score = keep.cpu().numpy(), num.cpu().numpy(), score.cpu().numpy()
    ok = True
    for i, c in enumerate(counts):
        if c == 0:
            ok &= num[i] == 0
            continue
        want_s = onms.rescore(ks[i], scs[i], rvis)
        ok &= np.array_equal(score[off[i]:off[i] + c], want_s)
        flat = ks[i].reshape(c, 51)
        want = (onms.soft_oks_nms(flat, ars[i], want_s, thr, 20, vis_thr=vthr) if soft
                else onms.oks_nms(flat, ars[i], want_s, thr, vis_thr=vthr))
        ok &= num[i] == len(want) and np.array_equal(keep[off[i]:off[i] + num[i]], want)
    print(f"case {it:3d} counts={counts} thr={thr} soft={soft} iou_vis={vthr}: {'ok' if ok else 'MISMATCH'}",
          flush=True)
    bad += 0 if ok else 1
print("mismatches:", bad)
sys.exit(1 if bad else 0)
