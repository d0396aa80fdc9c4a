#!/usr/bin/env python
"""Where the time of the keypoint gather goes (run under torchrun, N >= 2): CUDA-event times of
the scatter kernel, the flag wait, the blocking form and the NCCL all-gather, 4096 rows each.
Development aid."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mindpose_b200 import dist as pdist  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n, k = 4096, 17
preds = torch.rand(n, k, 3, device=dev)
boxes = torch.rand(n, 6, device=dev)
g = pdist.PeerGather(n, k, dev)
stream = torch.cuda.current_stream()


def timed(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(iters):
        fn()
    b.record(stream)
    b.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


res = {}
last = [None]


def scatter_only():
    last[0] = g.gather_async(preds, boxes)


res["scatter_signal_us"] = timed(scatter_only) * 1e3
last[0].wait()
torch.cuda.synchronize()
res["blocking_gather_us"] = timed(lambda: g.gather(preds, boxes)) * 1e3
prev = [None]


def deferred():
    t = g.gather_async(preds, boxes)
    if prev[0] is not None:
        prev[0].wait()
    prev[0] = t


res["deferred_gather_us"] = timed(deferred) * 1e3
prev[0].wait()
res["nccl_all_gather_us"] = timed(lambda: pdist.all_gather_keypoints(preds, boxes, world * n)) * 1e3
res["multicast"] = g.multicast


def graphed(label, body, reps=30):
    """device time per step of `body` (scatter [+ wait]) from a captured graph of `reps` steps"""
    torch.cuda.synchronize()
    dist.barrier()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            body()
    g.captured(reps)
    gr.replay()
    g.replayed(reps)
    g.wait_lag(0)
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(5):
        gr.replay()
        g.replayed(reps)
    g.wait_lag(0)
    b.record(stream)
    b.synchronize()
    t = torch.tensor([a.elapsed_time(b) / (5 * reps)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[label] = float(t) * 1e3


def body_deferred():
    g.gather_async(preds, boxes)
    g.wait_lag(1)


def body_blocking():
    g.gather_async(preds, boxes)
    g.wait_lag(0)


graphed("graph_deferred_us", body_deferred)
graphed("graph_blocking_us", body_blocking)
g.multicast = False
graphed("graph_deferred_peer_ptr_us", body_deferred)
g.multicast = res["multicast"]
# without multicast (peer pointers)
g.multicast = False
res["scatter_signal_peer_ptr_us"] = timed(scatter_only) * 1e3
last[0].wait()
torch.cuda.synchronize()
if rank == 0:
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
