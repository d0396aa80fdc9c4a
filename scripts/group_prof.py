"""Time group_by_tag on the config-4 inputs of bench.py (and on a crowded variant) with the
library POSECODEC_LIB selects; an experiment build (-DPC_GROUP_PROFILE=1) also prints the
kernel's cycles per phase.  Usage: POSECODEC_LIB=.../libposecodec_gprof.so python scripts/group_prof.py"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mindpose_b200 as mp  # noqa: E402
from mindpose_b200 import _lib, bottomup, synth  # noqa: E402

K = 17


def inputs(dev, people, seed=17):
    n = 64
    g = torch.Generator(device=dev).manual_seed(seed)
    out0 = torch.rand(n, 2 * K, 128, 128, device=dev, generator=g) * 0.02
    out1 = torch.rand(n, K, 256, 256, device=dev, generator=g) * 0.02
    ys = torch.randint(8, 248, (n, people), device=dev, generator=g)
    xs = torch.randint(8, 248, (n, people), device=dev, generator=g)
    ni = torch.arange(n, device=dev)[:, None, None]
    ki = torch.arange(K, device=dev)[None, :, None]
    jy = (ys[:, None, :] + torch.randint(-6, 7, (n, K, people), device=dev, generator=g)).clamp(2, 253)
    jx = (xs[:, None, :] + torch.randint(-6, 7, (n, K, people), device=dev, generator=g)).clamp(2, 253)
    out1[ni, ki, jy, jx] += 0.5 + 0.4 * torch.rand(n, K, people, device=dev, generator=g)
    out0[ni, ki, jy // 2, jx // 2] += 0.5
    tagv = (torch.arange(people, device=dev, dtype=torch.float32) * 3.0)[None, None, :].expand(n, K, people)
    out0[ni, ki + K, jy // 2, jx // 2] = tagv + 0.05 * torch.randn(n, K, people, device=dev, generator=g)
    mask = torch.ones(n, 512, 512, dtype=torch.uint8, device=dev)
    dec = mp.create_decoder("bottomup_heatmap_ae", use_nms=True, nms_kernel=3, max_num=30)
    dec.return_maps = False
    return dec([out0, out1], mask)[:3]


def main():
    dev = torch.device("cuda:0")
    lib = _lib.load()
    prof = getattr(lib, "pc_group_profile", None)
    print("library:", _lib.LIB_PATH, "(profile build)" if prof else "")
    for people in (8, 20, -12):
        if people < 0:
            # clean tags: 12 people, every joint detected, tags 3 apart with little noise (no
            # fragments) -- every joint's assignment has a strict, distinct minimum per row
            g = torch.Generator(device=dev).manual_seed(5)
            n, m, p = 64, 30, -people
            val_k = torch.zeros(n, K, m, device=dev)
            tag_k = torch.zeros(n, K, m, 1, device=dev)
            ind_k = torch.randint(0, 256, (n, K, m, 2), device=dev, generator=g).float()
            val_k[:, :, :p] = 0.9 - 0.05 * torch.arange(p, device=dev)[None, None]
            tag_k[:, :, :p, 0] = 3.0 * torch.arange(p, device=dev)[None, None] + \
                0.1 * torch.randn(n, K, p, device=dev, generator=g)
        else:
            val_k, tag_k, ind_k = inputs(dev, people)
        order = synth.COCO_JOINT_ORDER
        for _ in range(5):
            _, num, _ = bottomup.group_by_tag(val_k, tag_k, ind_k, order)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        times = []
        for _ in range(30):
            ev[0].record()
            bottomup.group_by_tag(val_k, tag_k, ind_k, order)
            ev[1].record()
            torch.cuda.synchronize()
            times.append(ev[0].elapsed_time(ev[1]))
        times.sort()
        print(f"people {people}: groups/image mean {num.float().mean().item():.1f} max {int(num.max())}; "
              f"group_by_tag median {times[len(times) // 2]:.4f} ms  min {times[0]:.4f} ms")
        if prof:
            out = (ctypes.c_ulonglong * 8)()
            prof(out, 1)
            bottomup.group_by_tag(val_k, tag_k, ind_k, order)
            prof(out, 1)
            tot = sum(out) or 1
            names = ["load detections", "refs + cost matrix", "assignment", "apply pairs"]
            print("  cycles per image: " + ", ".join(
                f"{nm} {out[i] / 64:.0f} ({100 * out[i] / tot:.0f}%)" for i, nm in enumerate(names)))


if __name__ == "__main__":
    main()
