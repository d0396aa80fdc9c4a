#!/usr/bin/env python
"""bench.py -- heatmap-codec throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--crops C] [--upload full|roi|roi_kernel] [--nccl-gather]
                    [--no-e2e] [--no-cpu-baseline] [--no-configs]

One "step" is one pass of the hot path over one batch of synthetic input.

Headline (BASELINE.json configs[1], "HRNet-W32 256x192 top-down eval codec"): affine crop
warp of 4096 u8 480x640x3 source images to 256x192x3, then DARK-refined decode with flip
averaging of 4096 x 2 x [17,64,48] float32 heatmaps.  `value`: inputs resident in HBM
(5.5 GB per step >> 126 MB of L2, so every step streams from DRAM), CUDA events on the
launching stream, max over ranks.  `e2e`: the same step through the host-buffer C-ABI front
end with pinned host inputs and host results, next to the host<->device link rate measured in
the same process with every rank copying at once (`e2e.link_gbs`).

`configs`: the other four BASELINE configs, device-timed on the same box in the same run --
config 1 (SimpleBaseline batch 64: Gaussian encode + quarter-offset flip decode), config 3
(UDP 96x72 batch 2048: UDP encode + UDP decode), config 4 (HigherHRNet batch 64: bottom-up
decode + tag grouping, with the CPU port of match_by_tag timed beside it) and config 5 (1 M
crops in 65,536-crop chunks; at N > 1 the STRONG-scaling sweep: the million crops sharded
contiguously over the ranks, then one all-gather of the keypoints, timed separately).

Under torchrun every rank runs its own 4096 crops per step (weak scaling) and the decoded
keypoints are gathered inside the timed region (one kernel of peer stores per step; the wait
for the other ranks' rows is taken one step later).  The gathered table is verified once per
run against every rank's own rows.  Prints ONE JSON line (rank 0).
"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "hrnet_eval": dict(image_size=[192, 256], heatmap_size=[48, 64], src_hw=(480, 640),
                       decoder=dict(dark_udp_refine=True, kernel_size=11), shift_heatmap=False,
                       crops=4096, label="HRNet-W32 256x192 top-down eval codec: affine crop "
                       "warp + DARK decode with flip averaging"),
}
K = 17
SWEEP_CROPS, SWEEP_CHUNK = 1_000_000, 65536


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


def pin_to_gpu_numa_node(index: int):
    """Bind this process (and the threads it starts) to the CPUs NVML names as local to GPU
    `index`, so that pinned staging buffers are first touched on the GPU's own NUMA node.
    Returns a short description for the JSON line."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1]
        cpus = [c for c in cpus if c < (os.cpu_count() or 0)]
        if not cpus:
            return {"pinned": False, "why": "NVML reports no local CPUs"}
        os.sched_setaffinity(0, cpus)
        node = None
        try:
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            with open(f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node") as f:
                node = int(f.read())
        except Exception:
            pass
        return {"pinned": True, "cpus": len(cpus), "numa_node": node}
    except Exception as e:  # no NVML / no permission: run unpinned and say so
        return {"pinned": False, "why": f"{type(e).__name__}: {e}"[:120]}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_bytes(kernel):
    """dram__bytes_read + dram__bytes_write of one kernel per launch, from the committed
    `ncu --set full` capture of this bench's command -- but only while the kernel source is
    the one that was profiled: profiles/ncu_traffic.json (scripts/ncu_traffic.py) stores the
    SHA-256 of the .cu file next to the bytes, and a stale entry reads as null."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        ent = json.load(f).get(kernel)
    if not ent:
        return None, None
    src = os.path.join(ROOT, ent["cu"])
    if not os.path.exists(src):
        return None, None
    with open(src, "rb") as f:
        if hashlib.sha256(f.read()).hexdigest() != ent["cu_sha256"]:
            return None, f"{ent['source']} is older than {ent['cu']}"
    return int(ent["dram_bytes"]), ent["source"]


def device_timer(torch, stream):
    def timeit(fn, iters, warmup=3):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(iters):
            fn()
        b.record(stream)
        b.synchronize()
        return a.elapsed_time(b) / iters
    return timeit


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
    Returns (crops_per_s, ms_per_step, crops per step)."""
    import multiprocessing as mp

    from mindpose_b200 import synth

    hw, ww = wl["heatmap_size"][1], wl["heatmap_size"][0]
    per = max(1, min(128, sample_crops // cores))
    nchunks = max(1, sample_crops // per)
    images, boxes = synth.source_images_and_boxes(per, *wl["src_hw"], seed=0)
    hm, _ = synth.blob_heatmaps(per, K, hw, ww, seed=0)
    fl = synth.flipped_pair(hm, seed=0)
    center, scale, score = synth.crop_geometry(per, seed=0)
    job = (images, boxes, hm, fl, center, scale, score, wl)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(warmup):
            pool.map(_cpu_decode_chunk, [job] * nchunks)
        t0 = time.perf_counter()
        done = 0
        for _ in range(steps):
            done += sum(pool.map(_cpu_decode_chunk, [job] * nchunks))
        dt = time.perf_counter() - t0
    return done / dt, dt / steps * 1e3, per * nchunks


def headline_config(wl, n):
    """The `config` object of the JSON line, identical in both arms (the reference arm runs a
    bounded SAMPLE of this workload per step and says so in cpu_baseline.sample)."""
    return {"workload": wl["label"], "crops_per_step_per_gpu": n,
            "l2": "inputs larger than L2 (5.5 GB resident per step)"}


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = args.crops or wl["crops"]
    sample = min(n, max(cores * 16, 256))
    steps, warmup = max(1, args.steps), max(3, args.warmup)   # the same rule as the other arm
    value, ms, per_step = cpu_reference_run(wl, sample, steps, warmup, cores)
    line = {
        "impl": "reference",
        "metric": "person-crops/sec encode+decode", "value": value, "unit": "crops/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": headline_config(wl, n),
        "cpu_baseline": {"value": value, "unit": "crops/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} of the {n} crops per step x {steps} steps: "
                                   "cv2.warpAffine per crop + the numpy restatement of the "
                                   "MindSpore decoder (mindspore is not installable here), "
                                   "batches of <= 128, one process per host core"},
        "e2e": {"value": value, "unit": "crops/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------ the other configs
def blob_stack(torch, n, h, w, dev, gen, fidx=None):
    """n x K blob heat maps built on the device (+ the mirrored stack when fidx is given)."""
    cxy = torch.rand(n, K, 2, device=dev, generator=gen)
    cx = 3 + cxy[..., 0] * (w - 7)
    cy = 3 + cxy[..., 1] * (h - 7)
    amp = 0.3 + 0.7 * torch.rand(n, K, device=dev, generator=gen)
    xs = torch.arange(w, device=dev, dtype=torch.float32)
    ys = torch.arange(h, device=dev, dtype=torch.float32)
    heat = torch.empty(n, K, h, w, device=dev)
    flip = torch.empty(n, K, h, w, device=dev) if fidx is not None else None
    inv_fidx = torch.as_tensor(np.argsort(fidx), device=dev) if fidx is not None else None
    step_c = max(1, (64 << 20) // (K * h * w * 4))
    for i0 in range(0, n, step_c):
        sl = slice(i0, min(n, i0 + step_c))
        d2 = (xs[None, None, None, :] - cx[sl, :, None, None]) ** 2 + \
             (ys[None, None, :, None] - cy[sl, :, None, None]) ** 2
        blob = amp[sl, :, None, None] * torch.exp(-d2 / 8.0)
        heat[sl] = blob + 0.02 * torch.rand(blob.shape, device=dev, generator=gen)
        if flip is not None:
            flip[sl] = blob[:, inv_fidx].flip(-1) + 0.02 * torch.rand(blob.shape, device=dev,
                                                                      generator=gen)
    return heat, flip


def kernel_entry(name, ms, units, bytes_per_unit, peak):
    gbs = units * bytes_per_unit / (ms * 1e-3) / 1e9
    return {"kernel": name, "ms": ms, "gbs": gbs, "frac": gbs / peak,
            "algorithmic_bytes_per_unit": int(bytes_per_unit)}


def run_config1(torch, dev, timeit, peak, iters):
    """BASELINE configs[0]: SimpleBaseline 256x192, batch 64: Gaussian target encode +
    TopDownHeatMapDecoder (quarter-offset shift) with flip test.  64 crops are 27 MB (< L2), so
    the timed loop rotates over 32 distinct batches (855 MB of heat maps + 427 MB of targets)."""
    import mindpose_b200 as mp
    from mindpose_b200 import codec, synth

    n, h, w, sets = 64, 64, 48, 32
    g = torch.Generator(device=dev).manual_seed(11)
    heat, flip = blob_stack(torch, n * sets, h, w, dev, g, synth.flip_index())
    kps = torch.from_numpy(synth.keypoints(n * sets, K, [192, 256], seed=1)).to(dev)
    target = torch.empty(n * sets, K, h, w, device=dev)
    center = torch.rand(n * sets, 2, device=dev, generator=g) * 400
    scale = torch.rand(n * sets, 2, device=dev, generator=g) * 2.8 + 0.2
    score = torch.rand(n * sets, device=dev, generator=g)
    dec = mp.create_decoder("topdown_heatmap", shift_coordinate=True)
    p = dec._params(K, h, w, flip_index=synth.flip_index(), shift_heatmap=True)
    # pre-sliced views: the timed loops issue nothing but the two library calls
    sets_v = [(kps[i * n:(i + 1) * n], target[i * n:(i + 1) * n], heat[i * n:(i + 1) * n],
               flip[i * n:(i + 1) * n], center[i * n:(i + 1) * n], scale[i * n:(i + 1) * n],
               score[i * n:(i + 1) * n]) for i in range(sets)]

    def enc(v):
        codec.topdown_encode(v[0], [192, 256], [48, 64], sigma=2.0, out=v[1])

    def decd(v):
        codec.topdown_decode(v[2], v[4], v[5], v[6], flipped=v[3], params=p)

    def all_sets(fn):
        def run():
            for v in sets_v:
                fn(v)
        return run

    def graphed(run):
        """One CUDA graph of the 32 rotating launches: at batch 64 a Python-driven loop is
        bound by the host (about 15 us per call), the graph shows what the device does."""
        try:
            run()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                run()
            return gr.replay, "one CUDA graph of the 32 rotating launches"
        except Exception as e:  # capture refused: time the eager loop and say so
            torch.cuda.synchronize()
            return run, f"eager loop ({type(e).__name__}: capture refused)"

    it = max(3, iters // 4)
    f_e, how = graphed(all_sets(enc))
    f_d, _ = graphed(all_sets(decd))
    f_b, _ = graphed(all_sets(lambda v: (enc(v), decd(v))))
    ms_e, ms_d, ms = timeit(f_e, it) / sets, timeit(f_d, it) / sets, timeit(f_b, it) / sets
    # the same two kernels at a batch that fills the GPU (all 32 sets = 2048 crops at once)
    big_e = timeit(lambda: codec.topdown_encode(kps, [192, 256], [48, 64], sigma=2.0, out=target),
                   max(3, iters // 4))
    big_d = timeit(lambda: codec.topdown_decode(heat, center, scale, score, flipped=flip, params=p),
                   max(3, iters // 4))
    be, bd = K * h * w * 4 + 272, 2 * K * h * w * 4 + 228
    return {
        "workload": "SimpleBaseline ResNet-50 256x192: Gaussian target encode + quarter-offset "
                    "decode with flip test, batch 64 of 17x64x48",
        "units_per_step": n, "unit": "crops", "ms_per_step": ms, "value": n / (ms * 1e-3),
        "l2": "rotating over 32 distinct batches (1.3 GB) so no batch is L2-resident",
        "timing": how,
        "kernels": [kernel_entry("topdown_encode_kernel<gaussian>", ms_e, n, be, peak),
                    kernel_entry("topdown_decode_kernel<flip> quarter offset", ms_d, n, bd, peak)],
        "note": "27 MB per launch: launch- and latency-bound at batch 64 (about 4 us of HBM "
                "time); `saturated` is the same pair of kernels at 2048 crops per launch",
        "saturated": {"units_per_step": n * sets,
                      "kernels": [kernel_entry("topdown_encode_kernel<gaussian>", big_e, n * sets,
                                               be, peak),
                                  kernel_entry("topdown_decode_kernel<flip> quarter offset", big_d,
                                               n * sets, bd, peak)]},
    }


def run_config3(torch, dev, timeit, peak, iters):
    """BASELINE configs[2]: UDP HRNet-W48 384x288: UDP target encode + UDP decode (flip test)
    on 17x96x72 maps, batch 2048 (1.9 GB per stack)."""
    import mindpose_b200 as mp
    from mindpose_b200 import codec, synth

    n, h, w = 2048, 96, 72
    g = torch.Generator(device=dev).manual_seed(13)
    heat, flip = blob_stack(torch, n, h, w, dev, g, synth.flip_index())
    kps = torch.from_numpy(synth.keypoints(n, K, [288, 384], seed=3)).to(dev)
    target = torch.empty(n, K, h, w, device=dev)
    center = torch.rand(n, 2, device=dev, generator=g) * 400
    scale = torch.rand(n, 2, device=dev, generator=g) * 2.8 + 0.2
    score = torch.rand(n, device=dev, generator=g)
    # configs/udp/hrnet_w48_udp_ascend.yaml: decoder use_udp, kernel_size 11; target sigma 3
    dec = mp.create_decoder("topdown_heatmap", use_udp=True, dark_udp_refine=True, kernel_size=11)
    p = dec._params(K, h, w, flip_index=synth.flip_index(), shift_heatmap=False)

    def enc():
        codec.topdown_encode(kps, [288, 384], [72, 96], sigma=3.0, use_udp=True, out=target)

    def decd():
        codec.topdown_decode(heat, center, scale, score, flipped=flip, params=p)

    def both():
        enc()
        decd()

    ms_e, ms_d, ms = timeit(enc, iters), timeit(decd, iters), timeit(both, iters)
    return {
        "workload": "UDP HRNet-W48 384x288: UDP target encode + UDP (DARK-refined) decode with "
                    "flip test on 17x96x72 maps, batch 2048",
        "units_per_step": n, "unit": "crops", "ms_per_step": ms, "value": n / (ms * 1e-3),
        "l2": "inputs larger than L2 (3.8 GB of heat maps read, 1.9 GB of targets written)",
        "kernels": [kernel_entry("topdown_encode_kernel<udp>", ms_e, n, K * h * w * 4 + 272, peak),
                    kernel_entry("topdown_decode_kernel<flip> udp", ms_d, n,
                                 2 * K * h * w * 4 + 228, peak)],
    }


def _cpu_match_chunk(args):
    from oracle import grouping

    val, tag, ind, order = args
    for i in range(val.shape[0]):
        grouping.match_by_tag(val[i], tag[i], ind[i], order)
    return val.shape[0]


def run_config4(torch, dev, timeit, peak, iters, cpu_baseline):
    """BASELINE configs[3]: HigherHRNet-W32 512x512, batch 64: 128^2 + 256^2 aggregation,
    NMS + top-k (k = 30), tag grouping.  445 MB of network outputs per step (> L2)."""
    import mindpose_b200 as mp
    from mindpose_b200 import bottomup, synth

    n = 64
    g = torch.Generator(device=dev).manual_seed(17)
    out0 = torch.rand(n, 2 * K, 128, 128, device=dev, generator=g) * 0.02
    out1 = torch.rand(n, K, 256, 256, device=dev, generator=g) * 0.02
    people = 8
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
    val_k, tag_k, ind_k, _, _ = dec([out0, out1], mask)
    order = synth.COCO_JOINT_ORDER

    def decd():
        dec([out0, out1], mask)

    def grp():
        bottomup.group_by_tag(val_k, tag_k, ind_k, order)

    def both():
        v, t, i, _, _ = dec([out0, out1], mask)
        bottomup.group_by_tag(v, t, i, order)

    ms_d, ms_g, ms = timeit(decd, iters), timeit(grp, iters), timeit(both, iters)
    _, num, _ = bottomup.group_by_tag(val_k, tag_k, ind_k, order)

    def pipelined_graph(batches=8, replays=10):
        """The same work over consecutive batches as ONE CUDA graph with two branches: the
        grouping of batch b (64 warps, latency bound) on a second stream under the decode of
        batch b + 1 (HBM bound).  Every batch is decoded and grouped once; all outputs stay
        alive so that no buffer is shared between the branches.  Launched from Python the two
        streams are host bound (measured: 0.239 ms against 0.205 ms serial), hence the graph."""
        side = torch.cuda.Stream()
        graph = torch.cuda.CUDAGraph()
        keep = []
        torch.cuda.synchronize()
        with torch.cuda.graph(graph):
            cur = torch.cuda.current_stream()
            for _ in range(batches):
                v, t, i, _, _ = dec([out0, out1], mask)
                done = torch.cuda.Event()
                done.record(cur)
                side.wait_event(done)
                with torch.cuda.stream(side):
                    keep.append((v, t, i, bottomup.group_by_tag(v, t, i, order)))
            cur.wait_stream(side)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(replays):
            graph.replay()
        b.record()
        b.synchronize()
        ok = bool(torch.equal(keep[-1][3][1], num))   # same people per image as the serial run
        return a.elapsed_time(b) / (replays * batches), ok

    try:
        ms_p, ok_p = pipelined_graph()
        pipelined = {"ms_per_step": ms_p, "value": n / (ms_p * 1e-3), "same_result": ok_p,
                     "note": "8 consecutive batches as one CUDA graph, grouping of batch b on a "
                             "second branch under the decode of batch b + 1; `ms_per_step` / "
                             "`value` above are one batch from start to end"}
    except Exception as e:  # noqa: BLE001 -- a measurement extra must never take the line down
        torch.cuda.synchronize()
        pipelined = {"error": f"{type(e).__name__}: {e}"[:200]}

    # bytes the decode needs: both heat-map stacks once + the mask; the tag planes are only
    # gathered at the <= 30 kept positions per joint (not counted as read)
    need = K * 128 * 128 * 4 + K * 256 * 256 * 4 + 512 * 512 + 8160
    full = 2 * K * 128 * 128 * 4 + K * 256 * 256 * 4 + 512 * 512 + 8160
    ent = {
        "workload": "HigherHRNet-W32 512x512 bottom-up: 128x128 + 256x256 aggregation, top-k "
                    "NMS (k=30) and associative-embedding tag grouping, batch 64 images",
        "units_per_step": n, "unit": "images", "ms_per_step": ms, "value": n / (ms * 1e-3),
        "l2": "inputs larger than L2 (445 MB of network outputs per step)",
        "kernels": [dict(kernel_entry("bottomup_decode_pairs_kernel (+ mask rows)", ms_d, n, need,
                                      peak),
                         frac_counting_unread_tag_planes=full * n / (ms_d * 1e-3) / 1e9 / peak),
                    {"kernel": "group_by_tag_kernel", "ms": ms_g,
                     "note": "latency bound: one warp per image, 17 sequential assignment "
                             "problems; 8 KB per image"}],
        "people_per_image_mean": float(num.float().mean().item()),
        "pipelined": pipelined,
    }
    if cpu_baseline:
        import multiprocessing as mpr

        cores = os.cpu_count() or 1
        v, t, i = val_k.cpu().numpy(), tag_k.cpu().numpy(), ind_k.cpu().numpy()
        jobs = [(v[j:j + 4], t[j:j + 4], i[j:j + 4], order) for j in range(0, n, 4)]
        with mpr.get_context("fork").Pool(min(cores, len(jobs))) as pool:
            pool.map(_cpu_match_chunk, jobs[:2])
            t0 = time.perf_counter()
            reps = 0
            while time.perf_counter() - t0 < 3.0:
                pool.map(_cpu_match_chunk, jobs)
                reps += 1
            dt = time.perf_counter() - t0
        ent["cpu_match_by_tag"] = {
            "value": n * reps / dt, "unit": "images/s", "cores": min(cores, len(jobs)),
            "kind": "port", "gpu_value": n / (ms_g * 1e-3),
            "sample": f"{reps} x 64 images: numpy restatement of match_by_tag "
                      "(mindpose/utils/match.py:14-116) + scipy linear_sum_assignment"}
    return ent


def sweep_geometry(torch, idx):
    """centre / scale / score of crop i as a function of its GLOBAL index only."""
    f = idx.to(torch.float64)
    frac = lambda v: v - torch.floor(v)  # noqa: E731
    center = torch.stack([frac(f * 0.6180339887) * 400, frac(f * 0.7548776662) * 400], 1)
    scale = torch.stack([0.2 + frac(f * 0.5698402910) * 2.8, 0.2 + frac(f * 0.3819660113) * 2.8], 1)
    return center.float().contiguous(), scale.float().contiguous(), \
        frac(f * 0.2451223338).float().contiguous()


def run_config5(torch, dist, dev, peak, world, rank, nccl_only):
    """BASELINE configs[4]: 1 M crops of 17x64x48 (flip pair, DARK), sharded contiguously by
    crop index (mindpose/data/data_factory.py:59-66), streamed through each GPU in pieces of at
    most 65,536 crops, then ONE all-gather of the keypoints.  Crop i reads entry i mod 65,536 of
    one resident bank (27 GB) and takes its geometry from i alone, so every N decodes the same
    million crops: the checksum of the gathered table is the same at every N."""
    import mindpose_b200 as mp
    from mindpose_b200 import codec, synth
    from mindpose_b200 import dist as pdist

    h, w = 64, 48
    g = torch.Generator(device=dev).manual_seed(5)       # the SAME bank on every rank
    bank, bank_f = blob_stack(torch, SWEEP_CHUNK, h, w, dev, g, synth.flip_index())
    dec = mp.create_decoder("topdown_heatmap", dark_udp_refine=True, kernel_size=11)
    p = dec._params(K, h, w, flip_index=synth.flip_index(), shift_heatmap=False)
    lo, hi = pdist.shard_range(SWEEP_CROPS, rank, world)
    pieces = []
    c0 = lo
    while c0 < hi:
        c1 = min(hi, (c0 // SWEEP_CHUNK + 1) * SWEEP_CHUNK)
        pieces.append((c0, c1))
        c0 = c1
    center, scale, score = sweep_geometry(torch, torch.arange(lo, hi, device=dev))
    preds = torch.empty(hi - lo, K, 3, device=dev)
    boxes = torch.empty(hi - lo, 6, device=dev)
    stream = torch.cuda.current_stream()

    def decode_all():
        for c0, c1 in pieces:
            b0, m = c0 % SWEEP_CHUNK, c1 - c0
            a = c0 - lo
            codec.topdown_decode(bank[b0:b0 + m], center[a:a + m], scale[a:a + m], score[a:a + m],
                                 flipped=bank_f[b0:b0 + m], params=p,
                                 out=(preds[a:a + m], boxes[a:a + m]))

    decode_all()                                           # warm-up
    if world > 1:
        pdist.all_gather_keypoints(preds, boxes, SWEEP_CROPS)
        dist.barrier()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(stream)
    decode_all()
    e[1].record(stream)
    all_p, all_b = pdist.all_gather_keypoints(preds, boxes, SWEEP_CROPS) if world > 1 else (preds, boxes)
    e[2].record(stream)
    e[2].synchronize()
    t = torch.tensor([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[0].elapsed_time(e[2])],
                     device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dec_ms, gat_ms, tot_ms = (float(v) for v in t)
    per_crop = 2 * K * h * w * 4 + 228
    checksum = float(all_p.double().sum().item() + all_b.double().sum().item())
    ent = {
        "workload": "codec scaling sweep: 1M synthetic 17x64x48 crops sharded by image, flip "
                    "pair + DARK decode, all-gather of keypoints",
        "units_per_step": SWEEP_CROPS, "unit": "crops", "scaling": "strong",
        "chunk_crops": SWEEP_CHUNK, "crops_per_rank": -(-SWEEP_CROPS // world),
        "ms_per_step": tot_ms, "value": SWEEP_CROPS / (tot_ms * 1e-3),
        "decode_ms_max_over_ranks": dec_ms, "gather_ms_max_over_ranks": gat_ms,
        "value_decode_only": SWEEP_CROPS / (dec_ms * 1e-3),
        "gather": None if world == 1 else "nccl all_gather_into_tensor (228 MB table: bandwidth bound)",
        "gather_bytes_total": SWEEP_CROPS * 228 if world > 1 else 0,
        "l2": "inputs larger than L2 (27 GB bank streamed per 65,536 crops)",
        "kernels": [kernel_entry("topdown_decode_kernel<flip> dark", dec_ms,
                                 -(-SWEEP_CROPS // world), per_crop, peak)],
        "checksum_gathered_table": checksum,
    }
    del bank, bank_f
    return ent


# --------------------------------------------------------------------- our arm
def link_bandwidth(torch, dist, dev, world):
    """Pinned host <-> device copy rate with EVERY rank copying at once (1 GiB each way):
    the ceiling of the e2e number at this N.  -> per-rank GB/s (min over ranks) and the sum."""
    nbytes = 1 << 30
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = {}
    for name, (dst, src) in (("h2d", (d, h)), ("d2h", (h, d))):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        gbs = 3 * nbytes / (time.perf_counter() - t0) / 1e9
        t = torch.tensor([gbs, -gbs, gbs], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t[:1], op=dist.ReduceOp.SUM)
            dist.all_reduce(t[1:2], op=dist.ReduceOp.MAX)
        out[name] = {"sum_gbs": float(t[0]), "min_rank_gbs": float(-t[1])}
    del h, d
    return out


def run_ours(args, wl):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    affinity = pin_to_gpu_numa_node(local_rank)   # before torch starts its threads

    import torch
    import torch.distributed as dist

    import mindpose_b200 as mp
    from mindpose_b200 import _lib, codec, synth
    from mindpose_b200 import dist as pdist

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200; no CUDA device is visible")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    n = args.crops or wl["crops"]
    iw, ih = wl["image_size"]
    w, h = wl["heatmap_size"]
    hs, ws = wl["src_hw"]
    cfg = dict(synth.TOPDOWN_CONFIG, image_size=wl["image_size"], heatmap_size=wl["heatmap_size"])
    steps, warmup = max(1, args.steps), max(3, args.warmup)

    # ---- synthetic inputs (seeded per rank)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    images = torch.randint(0, 256, (n, hs, ws, 3), device=dev, dtype=torch.uint8, generator=g)
    rng = np.random.RandomState(rank)
    bw = rng.uniform(40, 400, n)
    bh = rng.uniform(60, 440, n)
    boxes_np = np.stack([rng.uniform(0, 1, n) * (ws - bw), rng.uniform(0, 1, n) * (hs - bh), bw, bh],
                        axis=1).astype(np.float32)
    boxes = torch.from_numpy(boxes_np).to(dev)
    fidx = synth.flip_index()
    heat, flip = blob_stack(torch, n, h, w, dev, g, fidx)
    score = torch.rand(n, device=dev, generator=g)

    box_t = mp.create_transform("topdown_box_to_center_scale", is_train=False, config=cfg)
    decoder = mp.create_decoder("topdown_heatmap", **wl["decoder"])
    dparams = decoder._params(K, h, w, flip_index=fidx, shift_heatmap=wl["shift_heatmap"])
    crops = torch.empty(n, ih, iw, 3, device=dev, dtype=torch.uint8)
    off = torch.arange(n, device=dev, dtype=torch.int64) * (hs * ws * 3)
    src_hw = torch.tensor([hs, ws], device=dev, dtype=torch.int32).repeat(n, 1).contiguous()

    stream = torch.cuda.current_stream()
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    marks = []
    gatherer = None
    if world > 1 and not args.nccl_gather:
        gatherer = pdist.make_gatherer(n, K, dev)

    # The scatter of step t and the wait for step t - 1 run on a SIDE stream, next to the
    # parameter kernels and the crop warp of step t + 1; the main stream joins before the next
    # decode, which overwrites the results the scatter reads.  (--fused-gather instead lets the
    # decode kernel store into every rank's table itself: one kernel, measured 11 us per step
    # slower at N = 8, because its stores are 4 bytes each.)
    side = torch.cuda.Stream(device=dev) if gatherer is not None else None
    joins = [None]     # event: the side stream has finished the previous step's gather work

    def step(record=False):
        if record:
            e0, e1, e2 = ev(), ev(), ev()
            e0.record(stream)
        center, scale = box_t.box_to_center_scale_batch(boxes)
        _, inv = codec.affine_matrices(center, scale, None, cfg["image_size"])
        codec.warp_affine(images, off, src_hw, inv, cfg["image_size"], out=crops)
        if record:
            e1.record(stream)
        cur = torch.cuda.current_stream()   # (the capture stream while a graph is recorded)
        if joins[0] is not None:
            cur.wait_event(joins[0])
            joins[0] = None
        preds, bxs = codec.topdown_decode(heat, center, scale, score, flipped=flip, params=dparams,
                                          gather=gatherer if args.fused_gather else None)
        if record:
            e2.record(stream)
            marks.append((e0, e1, e2))
        if world > 1:
            if gatherer is None:         # NCCL all-gather
                pdist.all_gather_keypoints(preds, bxs, world * n)
            elif args.fused_gather:
                gatherer.wait_lag(1)
            else:
                done = torch.cuda.Event()
                done.record(cur)
                side.wait_event(done)
                with torch.cuda.stream(side):
                    gatherer.gather_async(preds, bxs)   # one kernel of peer stores + a flag
                    # the arrival of the PREVIOUS step's rows (this step's: in the next step)
                    gatherer.wait_lag(1)
                    joins[0] = torch.cuda.Event()
                    joins[0].record(side)
        return preds, bxs

    def drain():
        if joins[0] is not None:
            torch.cuda.current_stream().wait_event(joins[0])
            joins[0] = None
        if gatherer is not None:
            gatherer.wait_lag(0)

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    drain()
    fence()

    # ---- the gathered table, checked once: every rank holds every rank's rows bit for bit
    gather_verified = None
    if world > 1:
        preds, bxs = step()
        drain()
        got = gatherer.table_of_lag(0) if gatherer is not None else \
            pdist.all_gather_keypoints(preds, bxs, world * n)
        i64sum = lambda t: t.contiguous().view(torch.int32).to(torch.int64).sum()  # noqa: E731
        mine = torch.stack([i64sum(preds), i64sum(bxs)])
        owners = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(owners, mine)
        okv = 1
        for r in range(world):
            blk = torch.stack([i64sum(got[0][r * n:(r + 1) * n]), i64sum(got[1][r * n:(r + 1) * n])])
            okv &= int(torch.equal(blk, owners[r]))
        okt = torch.tensor([okv], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        gather_verified = bool(okt.item())
        fence()

    # ---- the split of a step between its kernels: a few eager steps with events inside
    for _ in range(12):
        step(record=True)
    drain()
    fence()
    warp_ms = float(np.median([a.elapsed_time(b) for a, b, _ in marks[2:]]))
    dec_ms = float(np.median([b.elapsed_time(c) for _, b, c in marks[2:]]))

    # ---- the timed steps: replays of ONE captured CUDA graph of the step (three steps when
    # the peer gather rotates its three tables), so that the host's launch work -- about as
    # long as the kernels themselves -- is not what is measured; NCCL gathers run eagerly
    unroll = pdist.PeerGather.TABLES if gatherer is not None else 1
    graph, how = None, "eager launches"
    if not args.no_graph and (world == 1 or gatherer is not None):
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for _ in range(unroll):
                    step()
                if joins[0] is not None:      # a capture must end with every fork joined
                    torch.cuda.current_stream().wait_event(joins[0])
                    joins[0] = None
            how = f"CUDA graph of {unroll} step(s), replayed"
            if gatherer is not None:
                gatherer.captured(unroll)  # the capture itself ran nothing
                graph.replay()             # warm the graph once (and check it runs)
                gatherer.replayed(unroll)
            else:
                graph.replay()
            drain()
            fence()
        except Exception as e:           # capture refused: fall back to eager launches
            graph, how = None, f"eager launches ({type(e).__name__}: graph capture refused)"
            torch.cuda.synchronize()
    q, r = (steps // unroll, steps % unroll) if graph is not None else (0, steps)
    t_start, t_end = ev(), ev()
    dbg = []
    with ClockSampler(local_rank) as clocks:   # (its thread starts before the ranks line up)
        fence()
        # One untimed unit lines the ranks up ON THE DEVICE: every step waits for every rank's
        # previous one, so after it the ranks are within a step of each other  (The host
        # leaves the barrier above up to 2 ms apart on an 8-GPU box, measured; a region of
        # 20 x 0.7 ms would otherwise time that skew, not the steps.)
        if graph is not None:
            graph.replay()
            if gatherer is not None:
                gatherer.replayed(unroll)
        else:
            step()
        drain()   # ... and within microseconds once each has seen every rank's last rows
        t_start.record(stream)
        for _ in range(q):
            graph.replay()
            if gatherer is not None:
                gatherer.replayed(unroll)
            if args.debug_steps:
                dbg.append(ev())
                dbg[-1].record(stream)
        for _ in range(r):
            step()
            if args.debug_steps:
                dbg.append(ev())
                dbg[-1].record(stream)
        drain()
        t_end.record(stream)
        fence()
    total_ms = t_start.elapsed_time(t_end)
    if args.debug_steps:
        marks_ms = [t_start.elapsed_time(e) for e in dbg] + [total_ms]
        print(f"rank {rank}: cumulative ms after each launch unit {[round(m, 3) for m in marks_ms]}",
              file=sys.stderr, flush=True)
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / steps
    value = world * n / (ms_per_step * 1e-3)

    # ---- e2e: the same step through the host-buffer front end (pinned host memory)
    e2e = None
    if not args.no_e2e:
        link = link_bandwidth(torch, dist, dev, world)
        hctx = codec.HostContext(local_rank, scratch_bytes=2 << 30)
        hctx2 = None if args.e2e_serial else codec.HostContext(local_rank, scratch_bytes=2 << 30)
        pin = lambda t: t.cpu().pin_memory()  # noqa: E731
        h_images, h_heat, h_flip = pin(images), pin(heat), pin(flip)
        h_boxes, h_score = boxes.cpu().numpy(), score.cpu().numpy()
        h_crops = torch.empty(n, ih, iw, 3, dtype=torch.uint8).pin_memory()
        h_preds = torch.empty(n, K, 3).pin_memory()
        h_bxs = torch.empty(n, 6).pin_memory()
        moved = [0, 0]   # bytes the library moved host->device / device->host in one step

        def warp_part():
            _, c_h, s_h = hctx.topdown_affine(h_images.numpy(), h_boxes, cfg["image_size"],
                                              out=h_crops.numpy(), upload=args.upload)
            return c_h, s_h, hctx.last_transfer_bytes()

        def decode_part(ctx, c_h, s_h):
            ctx.topdown_decode(h_heat.numpy(), c_h, s_h, h_score, flipped=h_flip.numpy(),
                               params=dparams, out_preds=h_preds.numpy(), out_boxes=h_bxs.numpy())
            return ctx.last_transfer_bytes()

        e2e_steps = max(1, min(steps, 10))
        if args.e2e_serial:
            def e2e_step():
                c_h, s_h, a = warp_part()
                b = decode_part(hctx, c_h, s_h)
                moved[0], moved[1] = a[0] + b[0], a[1] + b[1]

            e2e_step()
            fence()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            pipeline = "one call after the other"
        else:
            # The two calls of a step are different stages of the reference's pipeline (the
            # data loader warps, the inferencer decodes after the network), so in steady state
            # the crop warp of batch t + 1 and the decode of batch t are in flight together:
            # two contexts, two host threads (ctypes releases the GIL), one call of each per
            # step.  PCIe then never idles while one call drains its last chunk.
            from concurrent.futures import ThreadPoolExecutor

            pool = ThreadPoolExecutor(2)
            c_h, s_h, a = warp_part()                    # batch 0 (untimed prologue)
            decode_part(hctx2, c_h, s_h)
            fence()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                fw = pool.submit(warp_part)              # batch t + 1
                fd = pool.submit(decode_part, hctx2, c_h, s_h)   # batch t
                c_h, s_h, a = fw.result()
                b = fd.result()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            moved[0], moved[1] = a[0] + b[0], a[1] + b[1]
            pool.shutdown()
            pipeline = ("crop warp of batch t+1 and decode of batch t in flight together "
                        "(two contexts, two host threads; one call of each per step)")
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        # counted by the library from the copies it issued (pc_ctx_last_transfer_bytes); with
        # the default upload only the source rectangle each crop samples crosses PCIe
        step_s = dt / e2e_steps
        per_rank_gbs = (moved[0] + moved[1]) / step_s / 1e9
        e2e = {"value": world * n * e2e_steps / dt, "unit": "crops/s",
               "h2d_bytes_per_step": int(moved[0]), "d2h_bytes_per_step": int(moved[1]),
               "steps": e2e_steps, "ms_per_step": step_s * 1e3,
               "upload": args.upload or codec.DEFAULT_UPLOAD, "pipeline": pipeline,
               "h2d_bytes_per_step_whole_images": int(
                   images.numel() + 2 * heat.numel() * 4 + n * (16 + 8 + 8 + 4 + 8 + 8)),
               # the link: pinned copies of 1 GiB with all ranks copying at once, same process
               "link_gbs": link,
               "achieved_gbs_per_rank": per_rank_gbs,
               "link_frac": (moved[0] / step_s / 1e9) / max(link["h2d"]["min_rank_gbs"], 1e-9),
               "host_buffers": "pinned before the timed region (a pageable caller takes the "
                               "`roi` copy path, about 1.3x slower per step)",
               "cpu_affinity": affinity}
        hctx.close()
        if hctx2 is not None:
            hctx2.close()
        del h_images, h_heat, h_flip, h_crops

    # ---- the other BASELINE configs, device-timed
    configs = None
    if not args.no_configs:
        del images, crops
        torch.cuda.empty_cache()
        peak, _ = measured_peak_gbs()
        timeit = device_timer(torch, stream)

        def over_ranks(ent):
            """weak scaling of a per-rank config: max ms over ranks, N x the units"""
            if world > 1:
                t = torch.tensor([ent["ms_per_step"]], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ent["ms_per_step"] = float(t.item())
                ent["value"] = world * ent["units_per_step"] / (ent["ms_per_step"] * 1e-3)
                ent["scaling"] = "weak"
            return ent

        iters = max(10, steps)
        configs = {}
        configs["config1_simplebaseline_b64"] = over_ranks(run_config1(torch, dev, timeit, peak, iters))
        torch.cuda.empty_cache()
        configs["config3_udp_384_b2048"] = over_ranks(run_config3(torch, dev, timeit, peak, iters))
        torch.cuda.empty_cache()
        configs["config4_higherhrnet_b64"] = over_ranks(
            run_config4(torch, dev, timeit, peak, iters,
                        cpu_baseline=(rank == 0 and not args.no_cpu_baseline)))
        del heat, flip
        torch.cuda.empty_cache()
        configs["config5_sweep_1m"] = run_config5(torch, dist, dev, peak, world, rank,
                                                  args.nccl_gather)

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
                         n * (2 * K * h * w * 4 + (K * 3 + 6) * 4), dec_ms,
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
        r_warp = roofline("warp", "warp_affine_u8x3_band_kernel", n * ih * iw * 3 + roi, warp_ms,
                          "source rectangle staged in shared memory by bulk copies, taps "
                          "gathered from there: bound by the shared-memory data pipe (77 %) "
                          "ahead of HBM (63 %) in the committed ncu report; timed together with "
                          "the two parameter kernels")
        dominant, other = (r_warp, r_dec) if warp_ms >= dec_ms else (r_dec, r_warp)
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            v, _, per_step = cpu_reference_run(wl, max(cores * 16, 256), 2, 1, cores)
            cpu = {"value": v, "unit": "crops/s", "cores": cores, "kind": "port",
                   "sample": f"{per_step} crops per step x 2 steps (cv2 warp + numpy decode)"}
        launches = 4 + ((1 if args.fused_gather else 2) if gatherer is not None else 0)
        line = {
            "metric": "person-crops/sec encode+decode", "value": value, "unit": "crops/s",
            "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": headline_config(wl, n),
            "kernels_ms": {"warp": warp_ms, "decode": dec_ms},
            "launch": how,
            # the kernel with the largest share of the timed step, then the other one
            "roofline": dominant,
            "roofline_other": other,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": launches * steps,
            "gather": (None if world == 1 else
                       (("decode kernel stores into every rank's table" if args.fused_gather
                         else "scatter kernel of peer stores on a side stream") +
                        (" (NVSwitch multicast)" if gatherer.multicast else " (peer pointers)") +
                        " + flags, wait deferred by one step"
                        if gatherer is not None else "nccl all_gather")),
            "gather_verified": gather_verified,
            "parity": "numpy/cv2/scipy half bit-exact vs goldens made by the unmodified "
                      "reference; MindSpore half (decoders) vs an unpinned restatement",
            "configs": configs,
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
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-serial", action="store_true",
                    help="e2e: the two host-buffer calls of a step one after the other instead "
                         "of software-pipelined over two contexts")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the device-timed entries of BASELINE configs 1, 3, 4, 5")
    ap.add_argument("--upload", default=None, choices=["full", "roi", "roi_kernel"],
                    help="e2e: how the source images cross PCIe (default: codec.DEFAULT_UPLOAD)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--debug-steps", action="store_true",
                    help="print, per rank, the device time after every replay / eager step")
    ap.add_argument("--no-graph", action="store_true",
                    help="launch every timed step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--fused-gather", action="store_true",
                    help="N > 1: let the decode kernel store its results into every rank's "
                         "table itself (pc_topdown_decode_gather) instead of a scatter kernel "
                         "on a side stream")
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
