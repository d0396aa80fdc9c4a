#!/usr/bin/env python
"""bench.py -- heatmap-codec throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--crops C] [--upload full|roi|roi_kernel] [--nccl-gather]

(the other BASELINE configs are parity-test cases and rows of scripts/kbench.py, not bench lines)

One "step" is one pass of the hot path over one batch of synthetic input.
Default workload (BASELINE.json configs[1], "HRNet-W32 256x192 top-down eval
codec"): affine crop warp of 4096 u8 480x640x3 source images to 256x192x3, then
DARK-refined decode with flip averaging of 4096 x 2 x [17,64,48] float32
heatmaps.  Inputs are resident in HBM for `value` (5.5 GB per step >> 126 MB of
L2, so every step streams from DRAM); `e2e` repeats the step through the
host-buffer C-ABI front end with pinned host inputs and host results (of the
source images only the rectangle each crop samples is fetched over PCIe).

Prints ONE JSON line (rank 0).  Under torchrun each rank processes its own
4096 crops (weak scaling) and the decoded keypoints are all-gathered with NCCL
inside the timed region.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (image_size [w,h], heatmap [w,h], src hw, decoder kwargs, shift_heatmap, warp?)
    "hrnet_eval": dict(image_size=[192, 256], heatmap_size=[48, 64], src_hw=(480, 640),
                       decoder=dict(dark_udp_refine=True, kernel_size=11), shift_heatmap=False,
                       crops=4096, label="HRNet-W32 256x192 top-down eval codec: affine crop "
                       "warp + DARK decode with flip averaging"),
}


# ------------------------------------------------------------------ utilities
class ClockSampler:
    """SM clock / throttle reasons sampled with NVML while the timed region runs."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def ncu_traffic_bytes(kernel="decode"):
    """dram__bytes_read + dram__bytes_write of one kernel, per launch, from the latest
    committed `ncu --set full` summary under profiles/ (same workload as this bench)."""
    import glob
    import re

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"*_{kernel}_ncu_summary.txt")))
    if not files:
        return None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total = 0.0
    with open(files[-1]) as f:
        for line in f:
            m = re.search(r"dram__bytes_(read|write)\.sum = ([0-9.]+) (\w+)", line)
            if m:
                total += float(m.group(2)) * unit.get(m.group(3), 1.0)
    return (int(total) if total else None), os.path.basename(files[-1])


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------- reference arm
def _cpu_decode_chunk(args):
    """One chunk of the reference CPU path: cv2 warp per crop + numpy decode per
    batch of <= 128 crops (the reference's batch_size)."""
    import cv2

    from oracle import affine, topdown_decode

    cv2.setNumThreads(2)  # the reference sets this (topdown_transform.py:29)
    images, boxes, hm, fl, center, scale, score, wl = args
    image_size = np.array(wl["image_size"])
    for img, box in zip(images, boxes):
        c, s = affine.box_to_center_scale(tuple(box), image_size)
        m = affine.affine_matrix(c, s, 0.0, image_size)
        cv2.warpAffine(img, m, (int(image_size[0]), int(image_size[1])), flags=cv2.INTER_LINEAR)
    from mindpose_b200 import synth

    kw = dict(wl["decoder"])
    topdown_decode.decode_with_flip(
        hm, fl, synth.flip_index(), center, scale, score, shift_heatmap=wl["shift_heatmap"],
        dark_udp_refine_flag=kw.get("dark_udp_refine", False),
        shift_coordinate_flag=kw.get("shift_coordinate", False),
        use_udp=kw.get("use_udp", False), kernel_size=kw.get("kernel_size", 11))
    return len(images)


def cpu_reference_run(wl, sample_crops, steps, warmup, cores):
    """Times the reference CPU path (oracle port + cv2) on `cores` processes.
    Returns (crops_per_s, ms_per_step)."""
    import multiprocessing as mp

    from mindpose_b200 import synth

    hw, ww = wl["heatmap_size"][1], wl["heatmap_size"][0]
    per = max(1, min(128, sample_crops // cores))
    nchunks = max(1, sample_crops // per)
    images, boxes = synth.source_images_and_boxes(per, *wl["src_hw"], seed=0)
    hm, _ = synth.blob_heatmaps(per, 17, hw, ww, seed=0)
    fl = synth.flipped_pair(hm, seed=0)
    center, scale, score = synth.crop_geometry(per, seed=0)
    job = (images, boxes, hm, fl, center, scale, score, wl)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(warmup):
            pool.map(_cpu_decode_chunk, [job] * min(nchunks, cores))
        t0 = time.perf_counter()
        done = 0
        for _ in range(steps):
            done += sum(pool.map(_cpu_decode_chunk, [job] * nchunks))
        dt = time.perf_counter() - t0
    return done / dt, dt / steps * 1e3, per * nchunks


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = min(wl["crops"], max(cores * 16, 256))
    steps = max(1, args.steps)
    value, ms, per_step = cpu_reference_run(wl, sample, steps, min(args.warmup, 1), cores)
    line = {
        "impl": "reference",
        "metric": "person-crops/sec encode+decode", "value": value, "unit": "crops/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["label"], "crops_per_step": per_step,
                   "note": "reference CPU path: cv2.warpAffine per crop + numpy restatement of "
                           "the MindSpore decoder (mindspore not installable), batches of <=128"},
        "cpu_baseline": {"value": value, "unit": "crops/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} crops per step x {steps} steps"},
        "e2e": {"value": value, "unit": "crops/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------- our arm
def run_ours(args, wl):
    import torch
    import torch.distributed as dist

    import mindpose_b200 as mp
    from mindpose_b200 import _lib, codec, synth
    from mindpose_b200 import dist as pdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200; no CUDA device is visible")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    n = args.crops or wl["crops"]
    k = 17
    iw, ih = wl["image_size"]
    w, h = wl["heatmap_size"]
    hs, ws = wl["src_hw"]
    cfg = dict(synth.TOPDOWN_CONFIG, image_size=wl["image_size"], heatmap_size=wl["heatmap_size"])

    # ---- synthetic inputs, keyed by global crop index so every N sees the same data
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    images = torch.randint(0, 256, (n, hs, ws, 3), device=dev, dtype=torch.uint8, generator=g)
    _, boxes_np = synth.source_images_and_boxes(1, hs, ws, seed=rank)  # shapes only
    rng = np.random.RandomState(rank)
    bw = rng.uniform(40, 400, n)
    bh = rng.uniform(60, 440, n)
    boxes_np = np.stack([rng.uniform(0, 1, n) * (ws - bw), rng.uniform(0, 1, n) * (hs - bh), bw, bh],
                        axis=1).astype(np.float32)
    boxes = torch.from_numpy(boxes_np).to(dev)
    # blob heatmaps built on the device (the numpy generator is too slow for 1.7 GB)
    cxy = torch.rand(n, k, 2, device=dev, generator=g)
    cx = 3 + cxy[..., 0] * (w - 7)
    cy = 3 + cxy[..., 1] * (h - 7)
    amp = 0.3 + 0.7 * torch.rand(n, k, device=dev, generator=g)
    xs = torch.arange(w, device=dev, dtype=torch.float32)
    ys = torch.arange(h, device=dev, dtype=torch.float32)
    heat = torch.empty(n, k, h, w, device=dev)
    flip = torch.empty(n, k, h, w, device=dev)
    fidx = synth.flip_index()
    inv_fidx = torch.as_tensor(np.argsort(fidx), device=dev)
    step_c = 512
    for i0 in range(0, n, step_c):
        sl = slice(i0, min(n, i0 + step_c))
        d2 = (xs[None, None, None, :] - cx[sl, :, None, None]) ** 2 + \
             (ys[None, None, :, None] - cy[sl, :, None, None]) ** 2
        blob = amp[sl, :, None, None] * torch.exp(-d2 / 8.0)
        heat[sl] = blob + 0.02 * torch.rand(blob.shape, device=dev, generator=g)
        flip[sl] = blob[:, inv_fidx].flip(-1) + 0.02 * torch.rand(blob.shape, device=dev, generator=g)
    score = torch.rand(n, device=dev, generator=g)

    box_t = mp.create_transform("topdown_box_to_center_scale", is_train=False, config=cfg)
    decoder = mp.create_decoder("topdown_heatmap", **wl["decoder"])
    dparams = decoder._params(k, h, w, flip_index=fidx, shift_heatmap=wl["shift_heatmap"])
    crops = torch.empty(n, ih, iw, 3, device=dev, dtype=torch.uint8)
    off = torch.arange(n, device=dev, dtype=torch.int64) * (hs * ws * 3)
    src_hw = torch.tensor([hs, ws], device=dev, dtype=torch.int32).repeat(n, 1).contiguous()

    stream = torch.cuda.current_stream()
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    marks = []
    gatherer = None
    if world > 1 and not args.nccl_gather:
        gatherer = pdist.make_gatherer(n, k, dev)

    def step(record):
        if record:
            e0, e1, e2 = ev(), ev(), ev()
            e0.record(stream)
        center, scale = box_t.box_to_center_scale_batch(boxes)
        _, inv = codec.affine_matrices(center, scale, None, cfg["image_size"])
        codec.warp_affine(images, off, src_hw, inv, cfg["image_size"], out=crops)
        if record:
            e1.record(stream)
        preds, bxs = codec.topdown_decode(heat, center, scale, score, flipped=flip, params=dparams)
        if record:
            e2.record(stream)
            marks.append((e0, e1, e2))
        if world > 1:
            if gatherer is not None:     # one kernel of peer stores + a barrier
                gatherer.gather(preds, bxs)
            else:                        # NCCL all-gather
                pdist.all_gather_keypoints(preds, bxs, world * n)
        return preds, bxs

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step(False)
    fence()
    t_start, t_end = ev(), ev()
    with ClockSampler(local_rank) as clocks:
        t_start.record(stream)
        for _ in range(args.steps):
            step(True)
        t_end.record(stream)
        fence()
    total_ms = t_start.elapsed_time(t_end)
    warp_ms = float(np.mean([a.elapsed_time(b) for a, b, _ in marks]))
    dec_ms = float(np.mean([b.elapsed_time(c) for _, b, c in marks]))
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3)

    # ---- e2e: the same step through the host-buffer front end (pinned host memory)
    e2e = None
    if not args.no_e2e:
        hctx = codec.HostContext(local_rank, scratch_bytes=2 << 30)
        pin = lambda t: t.cpu().pin_memory()  # noqa: E731
        h_images, h_heat, h_flip = pin(images), pin(heat), pin(flip)
        h_boxes, h_score = boxes.cpu().numpy(), score.cpu().numpy()
        h_crops = torch.empty(n, ih, iw, 3, dtype=torch.uint8).pin_memory()
        h_preds = torch.empty(n, k, 3).pin_memory()
        h_bxs = torch.empty(n, 6).pin_memory()

        moved = [0, 0]   # bytes the library moved host->device / device->host in one step

        def e2e_step():
            _, c_h, s_h = hctx.topdown_affine(h_images.numpy(), h_boxes, cfg["image_size"],
                                              out=h_crops.numpy(), upload=args.upload)
            a = hctx.last_transfer_bytes()
            hctx.topdown_decode(h_heat.numpy(), c_h, s_h, h_score, flipped=h_flip.numpy(),
                                params=dparams, out_preds=h_preds.numpy(),
                                out_boxes=h_bxs.numpy())
            b = hctx.last_transfer_bytes()
            moved[0], moved[1] = a[0] + b[0], a[1] + b[1]

        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        e2e_step()
        fence()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        # counted by the library from the copies it issued (pc_ctx_last_transfer_bytes); with
        # the default upload only the source rectangle each crop samples crosses PCIe
        e2e = {"value": world * n * e2e_steps / dt, "unit": "crops/s",
               "h2d_bytes_per_step": int(moved[0]), "d2h_bytes_per_step": int(moved[1]),
               "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
               "upload": args.upload or codec.DEFAULT_UPLOAD,
               "h2d_bytes_per_step_whole_images": int(
                   images.numel() + 2 * heat.numel() * 4 + n * (16 + 8 + 8 + 4 + 8 + 8))}
        hctx.close()

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        full = n == wl["crops"]

        def roofline(kernel, name, alg_bytes, ms, note):
            achieved = alg_bytes / (ms * 1e-3) / 1e9
            traffic, traffic_src = ncu_traffic_bytes(kernel) if full else (None, None)
            return {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                    "traffic": traffic, "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": int(alg_bytes), "ms_per_launch": ms,
                    "share_of_step": ms / (warp_ms + dec_ms), "note": note}

        # decode: both heat-map stacks read once + results: 417,792 + 228 B per crop
        r_dec = roofline("decode", "topdown_decode_kernel<flip>",
                         n * (2 * k * h * w * 4 + (k * 3 + 6) * 4), dec_ms,
                         "every heat-map byte crosses HBM once")
        # warp: the crop written + the source region it samples (box * 1.25 padding, clipped)
        bw0, bh0 = boxes_np[:, 2].astype(np.float64), boxes_np[:, 3].astype(np.float64)
        asp = iw / ih
        sc = np.stack([np.where(bw0 < asp * bh0, bh0 * asp, bw0),
                       np.where(bw0 > asp * bh0, bw0 / asp, bh0)], axis=1) * cfg["scale_padding"]
        cxy_np = boxes_np[:, :2] + boxes_np[:, 2:4] / 2
        x0 = np.clip(cxy_np[:, 0] - sc[:, 0] / 2, 0, ws)
        x1 = np.clip(cxy_np[:, 0] + sc[:, 0] / 2, 0, ws)
        y0 = np.clip(cxy_np[:, 1] - sc[:, 1] / 2, 0, hs)
        y1 = np.clip(cxy_np[:, 1] + sc[:, 1] / 2, 0, hs)
        roi = float(np.sum((x1 - x0) * (y1 - y0) * 3))
        r_warp = roofline("warp", "warp_affine_u8x3_kernel", n * ih * iw * 3 + roi, warp_ms,
                          "integer gather bound on the SM, not by HBM: L1 data pipe 78 % and "
                          "instruction issue 74 % of their sustained peaks in the committed ncu "
                          "report; timed together with the two parameter kernels")
        dominant, other = (r_warp, r_dec) if warp_ms >= dec_ms else (r_dec, r_warp)
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            v, _, per_step = cpu_reference_run(wl, max(cores * 16, 256), 2, 1, cores)
            cpu = {"value": v, "unit": "crops/s", "cores": cores, "kind": "port",
                   "sample": f"{per_step} crops per step x 2 steps (cv2 warp + numpy decode)"}
        line = {
            "metric": "person-crops/sec encode+decode", "value": value, "unit": "crops/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["label"], "crops_per_step_per_gpu": n,
                       "l2": "inputs larger than L2 (5.5 GB resident per step)",
                       "kernels_ms": {"warp": warp_ms, "decode": dec_ms}},
            # the kernel with the largest share of the timed step, then the other one
            "roofline": dominant,
            "roofline_other": other,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": (4 + (1 if gatherer is not None else 0)) * args.steps,
            "gather": (None if world == 1 else
                       ("peer stores, multicast" if gatherer is not None and gatherer.multicast
                        else "peer stores" if gatherer is not None else "nccl all_gather")),
            "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hrnet_eval", choices=sorted(WORKLOADS))
    ap.add_argument("--crops", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--upload", default=None, choices=["full", "roi", "roi_kernel"],
                    help="e2e: how the source images cross PCIe (default: codec.DEFAULT_UPLOAD)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nccl-gather", action="store_true",
                    help="N > 1: use the NCCL all-gather instead of the peer-memory stores")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
