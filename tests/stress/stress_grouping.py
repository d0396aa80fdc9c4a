"""Randomised parity stress of pc_group_by_tag against the oracle (scipy's own
linear_sum_assignment inside): joints 3..32, max_num 1..64, up to 60 people, tag ties,
duplicate keys, crowded tags, rounded / exact norm, ignore_too_much, group capacities.
Development aid: python tests/stress/stress_grouping.py [cases]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mindpose_b200 import bottomup  # noqa: E402
from oracle import grouping  # noqa: E402


def random_batch(rng, n, k, m, people_max, mode):
    val = np.zeros((n, k, m), np.float32)
    tag = np.zeros((n, k, m, 1), np.float32)
    ind = np.zeros((n, k, m, 2), np.float32)
    for i in range(n):
        people = int(rng.randint(0, people_max + 1))
        for j in range(k):
            rows = []
            for pid in rng.permutation(people):
                if rng.random_sample() < 0.25:
                    continue
                if mode == "people":
                    t = pid * 3.0 + rng.normal(0, 0.15)
                elif mode == "ties":       # integer tags, repeated: ties and equal keys
                    t = float(pid % 7) + (0.0 if rng.random_sample() < 0.7 else 0.5)
                elif mode == "crowded":
                    t = pid * 0.6 + rng.normal(0, 0.2)
                else:                      # "wide": nearly every detection its own person
                    t = rng.uniform(-300, 300)
                rows.append((rng.uniform(0.05, 1.0), t))
            for _ in range(int(rng.randint(0, 4))):
                rows.append((rng.uniform(0.0, 0.5),
                             float(rng.randint(0, 8)) if mode == "ties" else rng.uniform(-1, 40)))
            rows = sorted(rows, key=lambda r: -r[0])[:m]
            for d, (v, t) in enumerate(rows):
                val[i, j, d], tag[i, j, d, 0] = v, t
                ind[i, j, d] = rng.randint(0, 256, 2)
    return val, tag, ind


def run(cases, seed=11, verbose=True):
    dev = torch.device("cuda", 0)
    rng = np.random.RandomState(seed)
    bad = 0
    for it in range(cases):
        k = int(rng.choice([3, 8, 17, 17, 17, 26, 32]))
        m = int(rng.choice([1, 5, 17, 30, 30, 31, 32, 33, 48, 64]))
        mode = str(rng.choice(["people", "ties", "crowded", "wide"]))
        people_max = int(rng.choice([3, 13, 30, 60]))
        rounded = bool(rng.rand() < 0.6)
        ignore = bool(rng.rand() < 0.3)
        vis_thr = float(rng.choice([0.1, 0.1, 0.3, 0.0]))
        tag_thr = float(rng.choice([1.0, 1.0, 0.4, 2.5]))
        order = [int(x) for x in rng.permutation(k)]
        room = bottomup.max_group_capacity(k, m)
        cap = None if rng.rand() < 0.5 else room
        n = 4
        val, tag, ind = random_batch(rng, n, k, m, people_max, mode)
        t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
        ans, num, scores = bottomup.group_by_tag(t(val), t(tag), t(ind), order, vis_thr=vis_thr,
                                                 tag_thr=tag_thr, ignore_too_much=ignore,
                                                 use_rounded_norm=rounded, max_groups=cap)
        ans, num, scores = ans.cpu().numpy(), num.cpu().numpy(), scores.cpu().numpy()
        ok, groups = True, []
        for i in range(n):
            want = grouping.match_by_tag(val[i], tag[i], ind[i], order, vis_thr=vis_thr,
                                         tag_thr=tag_thr, ignore_too_much=ignore,
                                         use_rounded_norm=rounded)
            p = 0 if want.ndim == 1 else want.shape[0]
            groups.append(p)
            if p > ans.shape[1]:
                ok &= num[i] == -1          # flagged, never silently truncated
                continue
            ok &= num[i] == p
            if p and num[i] == p:
                ok &= np.array_equal(ans[i, :p], want)
                ok &= scores[i, :p].tolist() == [float(x) for x in grouping.instance_scores(want)]
        if verbose:
            print(f"case {it:3d} K={k} M={m} {mode} people<={people_max} rounded={rounded} "
                  f"ignore={ignore} vis={vis_thr} tag={tag_thr} cap={ans.shape[1]} groups={groups}: "
                  f"{'ok' if ok else 'MISMATCH'}", flush=True)
        bad += 0 if ok else 1
    return bad


if __name__ == "__main__":
    bad = run(int(sys.argv[1]) if len(sys.argv) > 1 else 200)
    print("mismatches:", bad)
    sys.exit(1 if bad else 0)
