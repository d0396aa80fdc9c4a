#!/usr/bin/env python
"""BASELINE.json configs[4], the codec scaling sweep, on N GPUs of one box:

    python scripts/sweep_dist.py [--crops 1000000] [--chunk 65536]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/sweep_dist.py

1 M synthetic crops of 17 x 64 x 48 (flip pair; the SimpleBaseline decode recipe of config 1:
quarter-pixel shift, shifted flip heat map), sharded CONTIGUOUSLY by crop index over the ranks
(`dist.shard_range`, ragged tails included), streamed through each GPU in chunks of 65,536 crops
(27 GB per flip-pair chunk; 1 M crops are 418 GB and do not fit one GPU), then ONE NCCL
all-gather of the decoded keypoints [total, 57] f32 for evaluation.

Crop i's heat maps depend on the GLOBAL index i only (a bank of 4096 seeded blob stacks,
crop i = bank entry i mod 4096, geometry from a counter-based formula), so every N decodes
the same million crops and the float64 checksum of the gathered table must be identical at
every N -- the size-independent parity property of this config.  The timed region is the
decode kernels (CUDA events around each chunk's launch; building a chunk from the bank is
outside it) and, separately, the gather; both as the max over ranks.

Development aid (not a bench.py line): prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--crops", type=int, default=1_000_000)
    ap.add_argument("--chunk", type=int, default=65536)
    ap.add_argument("--bank", type=int, default=4096)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import mindpose_b200 as mp
    from mindpose_b200 import codec, synth
    from mindpose_b200 import dist as pdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("sweep_dist.py needs a B200; no CUDA device is visible")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    k, h, w = 17, 64, 48
    total = args.crops
    lo, hi = pdist.shard_range(total, rank, world)

    # the bank: identical on every rank (seeded on the host)
    maps, _ = synth.blob_heatmaps(args.bank, k, h, w, seed=5)
    bank = torch.from_numpy(maps).to(dev)
    bank_f = torch.from_numpy(synth.flipped_pair(maps, seed=5)).to(dev)
    del maps
    dec = mp.create_decoder("topdown_heatmap", shift_coordinate=True)
    p = dec._params(k, h, w, flip_index=synth.flip_index(), shift_heatmap=True)

    def geometry(idx):   # counter-based: a function of the global crop index only
        f = idx.to(torch.float64)
        frac = lambda v: v - torch.floor(v)  # noqa: E731
        center = torch.stack([frac(f * 0.6180339887) * 400, frac(f * 0.7548776662) * 400], 1)
        scale = torch.stack([0.2 + frac(f * 0.5698402910) * 2.8,
                             0.2 + frac(f * 0.3819660113) * 2.8], 1)
        score = frac(f * 0.2451223338)
        return center.float(), scale.float(), score.float()

    stream = torch.cuda.current_stream()
    preds_all = torch.empty(hi - lo, k, 3, device=dev)
    boxes_all = torch.empty(hi - lo, 6, device=dev)
    decode_ms = 0.0
    for c0 in range(lo, hi, args.chunk):
        c1 = min(hi, c0 + args.chunk)
        idx = torch.arange(c0, c1, device=dev)
        sel = idx % args.bank
        hm, fl = bank.index_select(0, sel), bank_f.index_select(0, sel)
        center, scale, score = geometry(idx)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        pr, bx = codec.topdown_decode(hm, center, scale, score, flipped=fl, params=p)
        e1.record(stream)
        e1.synchronize()
        decode_ms += e0.elapsed_time(e1)
        preds_all[c0 - lo:c1 - lo] = pr
        boxes_all[c0 - lo:c1 - lo] = bx
        del hm, fl
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(stream)
    all_p, all_b = pdist.all_gather_keypoints(preds_all, boxes_all, total)
    g1.record(stream)
    g1.synchronize()
    gather_ms = g0.elapsed_time(g1)
    t = torch.tensor([decode_ms, gather_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    decode_ms, gather_ms = float(t[0]), float(t[1])
    if rank == 0:
        assert all_p.shape[0] == total and all_b.shape[0] == total
        per_crop = 2 * k * h * w * 4 + (k * 3 + 6) * 4
        per_rank = -(-total // world)
        print(json.dumps({
            "config": "codec scaling sweep: 1M crops 17x64x48, flip pair, quarter shift",
            "n_gpus": world, "crops": total, "chunk": args.chunk,
            "decode_ms_max_over_ranks": decode_ms, "gather_ms": gather_ms,
            "crops_per_s": total / ((decode_ms + gather_ms) * 1e-3),
            "crops_per_s_decode_only": total / (decode_ms * 1e-3),
            "gbs_per_gpu_decode": per_rank * per_crop / (decode_ms * 1e-3) / 1e9,
            "gather_bytes": total * (k * 3 + 6) * 4,
            "checksum": float(all_p.double().sum() + all_b.double().sum()),
            "checksum_abs": float(all_p.double().abs().sum()),
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
