"""Per-kernel timing table (CUDA events, inputs >> L2, 3 warm-up + N timed launches).

    python scripts/kbench.py [--iters 10] [--only decode,encode,warp,bottomup,group] [--json out]

Reports algorithmic GB/s (SURVEY.md section 8(d) bytes per unit) and the fraction of the
measured HBM copy peak (MEASURED_PEAKS.json).  Development aid; bench.py is the contract.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mindpose_b200 as mp  # noqa: E402
from mindpose_b200 import bottomup, codec, synth  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs]
    return float(np.median(ms)), float(np.min(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="decode,encode,warp,bottomup,group,bu_encode,refine,nms,sweep,rescale")
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    only = set(args.only.split(","))
    dev = torch.device("cuda", 0)
    pk = peak()
    rows = []

    def report(name, units, unit_name, bytes_per_unit, ms_med, ms_min):
        gbs = units * bytes_per_unit / (ms_med * 1e-3) / 1e9
        rows.append(dict(kernel=name, units=units, unit=unit_name, ms=ms_med, ms_min=ms_min,
                         per_s=units / (ms_med * 1e-3), gbs=gbs, frac=gbs / pk))
        print(f"{name:46s} {ms_med:8.3f} ms  {units / (ms_med * 1e-3) / 1e6:8.3f} M{unit_name}/s "
              f"{gbs:8.1f} GB/s  {gbs / pk:5.3f} of measured peak", flush=True)

    k = 17
    if "decode" in only:
        for (h, w, n) in ((64, 48, 4096), (96, 72, 2048)):
            hm = torch.rand(n, k, h, w, device=dev) * 0.02
            fl = torch.rand(n, k, h, w, device=dev) * 0.02
            ys = torch.randint(4, h - 4, (n, k), device=dev)
            xs = torch.randint(4, w - 4, (n, k), device=dev)
            ni = torch.arange(n, device=dev)[:, None]
            ki = torch.arange(k, device=dev)[None, :]
            for dy in range(-3, 4):
                for dx in range(-3, 4):
                    g = float(np.exp(-(dx * dx + dy * dy) / 8.0))
                    hm[ni, ki, ys + dy, xs + dx] += 0.8 * g
            center = torch.rand(n, 2, device=dev) * 400
            scale = torch.rand(n, 2, device=dev) * 2.8 + 0.2
            score = torch.rand(n, device=dev)
            for mode, kw in (("plain", {}), ("quarter", dict(shift_coordinate=True)),
                             ("dark", dict(dark_udp_refine=True))):
                dec = mp.create_decoder("topdown_heatmap", **kw)
                for flip in (False, True):
                    if flip:
                        p = dec._params(k, h, w, flip_index=synth.flip_index(), shift_heatmap=True)
                        fn = lambda: codec.topdown_decode(hm, center, scale, score, flipped=fl, params=p)  # noqa: E731
                    else:
                        p = dec._params(k, h, w)
                        fn = lambda: codec.topdown_decode(hm, center, scale, score, params=p)  # noqa: E731
                    med, mn = timeit(fn, args.iters)
                    bpu = k * h * w * 4 * (2 if flip else 1) + 228
                    report(f"topdown_decode {h}x{w} {mode}{' flip' if flip else ''}", n, "crops",
                           bpu, med, mn)
            del hm, fl
    if "sweep" in only:
        # BASELINE config 5 on one GPU: 1 M crops of 17x64x48, flip pair + DARK, streamed as 16
        # chunks of 65,536 crops (27 GB per chunk; the same synthetic chunk is decoded 16 times)
        n, h, w, chunks = 65536, 64, 48, 16
        hm = torch.rand(n, k, h, w, device=dev) * 0.02
        hm[:, :, 20:23, 20:23] += 0.5
        fl = torch.rand(n, k, h, w, device=dev) * 0.02
        center = torch.rand(n, 2, device=dev) * 400
        scale = torch.rand(n, 2, device=dev) * 2.8 + 0.2
        score = torch.rand(n, device=dev)
        dec = mp.create_decoder("topdown_heatmap", dark_udp_refine=True)
        p = dec._params(k, h, w, flip_index=synth.flip_index(), shift_heatmap=True)

        def sweep():
            for _ in range(chunks):
                codec.topdown_decode(hm, center, scale, score, flipped=fl, params=p)

        med, mn = timeit(sweep, max(3, args.iters // 3))
        report("topdown_decode sweep: 1M crops, dark flip, 16 x 65,536", n * chunks, "crops",
               2 * k * h * w * 4 + 228, med, mn)
        del hm, fl
    if "encode" in only:
        for cfg, n in ((synth.TOPDOWN_CONFIG, 8192), (synth.TOPDOWN_CONFIG_384, 4096)):
            w, h = cfg["heatmap_size"]
            kps = torch.from_numpy(synth.keypoints(n, k, cfg["image_size"], seed=0)).to(dev)
            out = torch.empty(n, k, h, w, device=dev)
            for udp in (False, True):
                fn = lambda: codec.topdown_encode(kps, cfg["image_size"], cfg["heatmap_size"],  # noqa: E731
                                                  sigma=2.0, use_udp=udp, out=out)
                med, mn = timeit(fn, args.iters)
                report(f"topdown_encode {h}x{w} {'udp' if udp else 'gaussian'}", n, "crops",
                       k * h * w * 4 + 272, med, mn)
            del out
    if "warp" in only:
        n, hs, ws = 4096, 480, 640
        images = torch.randint(0, 256, (n, hs, ws, 3), device=dev, dtype=torch.uint8)
        rng = np.random.RandomState(0)
        bw, bh = rng.uniform(40, 400, n), rng.uniform(60, 440, n)
        boxes = torch.from_numpy(np.stack([rng.uniform(0, 1, n) * (ws - bw),
                                           rng.uniform(0, 1, n) * (hs - bh), bw, bh], 1)
                                 .astype(np.float32)).to(dev)
        center, scale = codec.box_to_center_scale(boxes, [192, 256])
        _, inv = codec.affine_matrices(center, scale, None, [192, 256])
        off = torch.arange(n, device=dev, dtype=torch.int64) * (hs * ws * 3)
        hw = torch.tensor([hs, ws], device=dev, dtype=torch.int32).repeat(n, 1).contiguous()
        out = torch.empty(n, 256, 192, 3, device=dev, dtype=torch.uint8)
        fn = lambda: codec.warp_affine(images, off, hw, inv, [192, 256], out=out)  # noqa: E731
        med, mn = timeit(fn, args.iters)
        # source ROI = scale * 200 in each direction, clipped to the image
        s = scale.cpu().numpy() * 200.0
        c = center.cpu().numpy()
        x0 = np.clip(c[:, 0] - s[:, 0] / 2, 0, ws)
        x1 = np.clip(c[:, 0] + s[:, 0] / 2, 0, ws)
        y0 = np.clip(c[:, 1] - s[:, 1] / 2, 0, hs)
        y1 = np.clip(c[:, 1] + s[:, 1] / 2, 0, hs)
        roi = float(np.mean((x1 - x0) * (y1 - y0) * 3))
        report("warp_affine_u8 480x640 -> 256x192 (dst+ROI)", n, "crops", 147456 + roi, med, mn)
        report("warp_affine_u8 (dst bytes only)", n, "crops", 147456, med, mn)
        # training-time crops: every crop rotated (the quad path inside the band kernel)
        rot = torch.from_numpy(rng.uniform(-40, 40, n).astype(np.float32)).to(dev)
        _, inv_r = codec.affine_matrices(center, scale, rot, [192, 256])
        fn = lambda: codec.warp_affine(images, off, hw, inv_r, [192, 256], out=out)  # noqa: E731
        med, mn = timeit(fn, args.iters)
        report("warp_affine_u8 rotated +-40 deg (dst+ROI of rot 0)", n, "crops", 147456 + roi, med, mn)
        outf = torch.empty(n, 3, 256, 192, device=dev)
        m255 = (np.array([0.485, 0.456, 0.406]) * 255.0).tolist()
        s255 = (np.array([0.229, 0.224, 0.255]) * 255.0).tolist()
        fn = lambda: codec.warp_affine_normalized(images, off, hw, inv, [192, 256], m255, s255, out=outf)  # noqa: E731
        med, mn = timeit(fn, args.iters)
        report("warp + normalize + CHW f32 (dst+ROI)", n, "crops", 589824 + roi, med, mn)
        del outf
        del images, out
    if "bottomup" in only:
        n = 64
        g = torch.Generator(device=dev).manual_seed(0)
        out0 = torch.rand(n, 34, 128, 128, device=dev, generator=g) * 0.02
        out1 = torch.rand(n, 17, 256, 256, device=dev, generator=g) * 0.02
        ys = torch.randint(8, 248, (n, 17, 8), device=dev, generator=g)
        xs = torch.randint(8, 248, (n, 17, 8), device=dev, generator=g)
        for j in range(8):
            out1[torch.arange(n)[:, None], torch.arange(17)[None, :], ys[..., j], xs[..., j]] += 0.5
        mask = torch.ones(n, 512, 512, dtype=torch.uint8, device=dev)
        dec = mp.create_decoder("bottomup_heatmap_ae", use_nms=True, nms_kernel=3, max_num=30)
        dec.return_maps = False
        fn = lambda: dec([out0, out1], mask)  # noqa: E731
        med, mn = timeit(fn, args.iters)
        bottomup.decode_stats(reset=True)
        fn()
        print(f"  exact-pass planes in one call: {bottomup.decode_stats()} of {n * 17}")
        report("bottomup_decode 64 x (34x128^2 + 17x256^2)", n, "images", 6946816 + 8160, med, mn)
        if "group" in only:
            val_k, tag_k, ind_k, _, _ = dec([out0, out1], mask)
            fn = lambda: bottomup.group_by_tag(val_k, tag_k, ind_k, synth.COCO_JOINT_ORDER)  # noqa: E731
            med, mn = timeit(fn, args.iters)
            report("group_by_tag 64 images (latency bound)", n, "images", 8160, med, mn)
    if "refine" in only:
        n, k, h, w, people = 64, 17, 256, 256, 8
        g = torch.Generator(device=dev).manual_seed(1)
        heat = torch.rand(n, k, h, w, device=dev, generator=g)
        tagm = torch.rand(n, k, h, w, 1, device=dev, generator=g) * 10
        ans = torch.zeros(n, bottomup._lib.PC_MAX_GROUPS, k, 4, device=dev)
        ans[:, :people, :, 0] = torch.randint(0, w, (n, people, k), device=dev, generator=g).float()
        ans[:, :people, :, 1] = torch.randint(0, h, (n, people, k), device=dev, generator=g).float()
        ans[:, :people, :, 2] = (torch.rand(n, people, k, device=dev, generator=g) > 0.4).float() * 0.5
        num = torch.full((n,), people, dtype=torch.int32, device=dev)
        fn = lambda: bottomup.refine_missing(heat, tagm, ans.clone(), num)  # noqa: E731
        med, mn = timeit(fn, args.iters)
        report("refine_missing 64 images x 8 people (maps read once)", n, "images",
               2 * k * h * w * 4, med, mn)
    if "nms" in only:
        from mindpose_b200 import nms as dnms
        images, people = 512, 40
        ks, ars, scs = [], [], []
        for i in range(8):                      # 8 distinct images, tiled
            rng = np.random.RandomState(900 + i)
            base = rng.randint(0, 10, people)   # 10 poses, each with jittered near-duplicates
            centers, sizes = rng.uniform(50, 400, (10, 2)), rng.uniform(40, 160, 10)
            shapes = rng.uniform(-0.5, 0.5, (10, 17, 2))
            kk = np.zeros((people, 17, 3), np.float32)
            jit = rng.choice([0.01, 0.03, 0.08, 0.2], people)[:, None, None] * sizes[base][:, None, None]
            kk[..., :2] = centers[base][:, None] + shapes[base] * sizes[base][:, None, None] + \
                rng.normal(0, 1, (people, 17, 2)) * jit
            kk[..., 2] = rng.uniform(0, 1, (people, 17))
            ks.append(kk)
            ars.append((sizes[base] ** 2 * rng.uniform(0.8, 1.2, people)).astype(np.float32))
            scs.append(rng.uniform(0.05, 1.0, people).astype(np.float32))
        reps = images // 8
        kp = torch.from_numpy(np.concatenate(ks * reps)).to(dev)
        ar = torch.from_numpy(np.concatenate(ars * reps)).to(dev)
        sc0 = torch.from_numpy(np.concatenate(scs * reps)).to(dev)
        off = torch.arange(images + 1, dtype=torch.int32, device=dev) * people
        for soft in (False, True):
            fn = lambda: dnms.rescore_and_nms(kp, ar, sc0.clone(), off, people, oks_thr=0.9,  # noqa: E731
                                              rescore_vis_thr=0.2, soft=soft)
            med, mn = timeit(fn, args.iters)
            report(f"oks {'soft ' if soft else ''}nms {images} images x {people} people (latency bound)",
                   images, "images", people * (17 * 12 + 8), med, mn)
    if "bu_encode" in only:
        n, m = 64, 30
        sizes = [[128, 128], [256, 256]]
        rng = np.random.RandomState(0)
        kp = np.zeros((n, 2, m, 17, 3), np.float32)
        kp[:, 0, :, :, 0] = rng.uniform(-5, 133, (n, m, 17))
        kp[:, 0, :, :, 1] = rng.uniform(-5, 133, (n, m, 17))
        kp[:, 0, :, :, 2] = (rng.random_sample((n, m, 17)) < 0.3) * (np.arange(m)[None, :, None] < 8)
        kp[:, 1] = kp[:, 0] * np.array([2, 2, 1], np.float32)
        kpt = torch.from_numpy(kp).to(dev)
        fn = lambda: bottomup.encode_targets(kpt, sizes)  # noqa: E731
        med, mn = timeit(fn, args.iters)
        report("bottomup_encode 64 x 2 x 17 x 256^2 (8 people)", n, "images",
               2 * 17 * 256 * 256 * 4 + 2 * 30 * 17 * 8, med, mn)
    if "rescale" in only:
        # bottom-up evaluation preprocessing: 256 images 480 x 640 -> 683 x 512 in an 832 x 512
        # canvas + mask (source read once, canvas and mask written once)
        n, sh, sw, cw, ch = 256, 480, 640, 832, 512
        g = torch.Generator(device=dev).manual_seed(0)
        imgs = torch.randint(0, 256, (n, sh, sw, 3), device=dev, dtype=torch.uint8, generator=g)
        off = torch.arange(n, device=dev, dtype=torch.int64) * (sh * sw * 3)
        hw = torch.tensor([[sh, sw]] * n, dtype=torch.int32, device=dev)
        wh = torch.tensor([[683, 512]] * n, dtype=torch.int32, device=dev)
        fn = lambda: codec.rescale_pad(imgs, off, hw, wh, (cw, ch))  # noqa: E731
        med, mn = timeit(fn, args.iters)
        report("rescale + pad + mask 480x640 -> 683x512 in 832x512 (256 images)", n, "images",
               sh * sw * 3 + cw * ch * 4, med, mn)
        fn = lambda: codec.rescale_pad(imgs, off, hw, wh, (cw, ch), mean=[123.675, 116.28, 103.53],  # noqa: E731
                                       std=[58.395, 57.12, 57.375])
        med, mn = timeit(fn, args.iters)
        report("... + normalize + CHW f32 (256 images)", n, "images",
               sh * sw * 3 + cw * ch * 13, med, mn)
        try:   # the reference's own two calls on one host core, beside it
            import time

            import cv2
            cv2.setNumThreads(1)
            host = imgs[:16].cpu().numpy()
            t0 = time.perf_counter()
            for im in host:
                r = cv2.resize(im, (683, 512), interpolation=cv2.INTER_LINEAR)
                np.pad(r, ((0, ch - 512), (0, cw - 683), (0, 0)))
                m = np.zeros((ch, cw), np.uint8)
                m[:512, :683] = 1
            dt = (time.perf_counter() - t0) / len(host)
            print(f"  cv2.resize + np.pad + mask on one host core: {dt * 1e3:.3f} ms per image "
                  f"({1 / dt:.0f} images/s)")
        except ImportError:
            pass
    if args.json:
        with open(args.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
