"""Peer-memory keypoint gather (pc_scatter_results_signal + pc_wait_peer_flags over symmetric
memory) on 2 GPUs against the NCCL all-gather: the blocking form and the deferred form
(ticket of step t waited for in step t + 1, three alternating tables).  Skipped on boxes with
one GPU; bench.py verifies the gathered table at every N it runs at."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, rows, k, out_dir):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from mindpose_b200 import dist as pdist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        g = pdist.PeerGather(rows, k, dev)

        def inputs(step):
            idx = torch.arange(rank * rows, (rank + 1) * rows, dtype=torch.float32, device=dev)
            preds = (idx[:, None, None] * 10 + step +
                     torch.arange(k * 3, dtype=torch.float32, device=dev).reshape(1, k, 3))
            boxes = idx[:, None] * 100 + step + torch.arange(6, dtype=torch.float32, device=dev)[None]
            return preds.contiguous(), boxes.contiguous()

        for step in range(4):       # blocking form: all three tables, reuse of the first
            preds, boxes = inputs(step)
            got_p, got_b = g.gather(preds, boxes)
            want_p, want_b = pdist.all_gather_keypoints(preds, boxes, world * rows)
            torch.cuda.synchronize()
            assert torch.equal(got_p, want_p) and torch.equal(got_b, want_b), (rank, step)
        # deferred form: the ticket of step t is waited for after the scatter of step t + 1,
        # with rank 1 lagging so that rank 0 runs a step ahead
        prev, prev_want = None, None
        for step in range(4, 12):
            if rank == 1:
                torch.cuda._sleep(3_000_000)
            preds, boxes = inputs(step)
            want = pdist.all_gather_keypoints(preds, boxes, world * rows)
            ticket = g.gather_async(preds, boxes)
            if prev is not None:
                got_p, got_b = prev.wait()
                assert torch.equal(got_p, prev_want[0]) and torch.equal(got_b, prev_want[1]), \
                    (rank, step)
            prev, prev_want = ticket, want
        got_p, got_b = prev.wait()
        torch.cuda.synchronize()
        assert torch.equal(got_p, prev_want[0]) and torch.equal(got_b, prev_want[1])

        # decode and all-gather as ONE kernel (pc_topdown_decode_gather): the decode kernel
        # stores its results into every rank's table itself
        import mindpose_b200 as mp
        from mindpose_b200 import codec, synth

        gen = torch.Generator(device=dev).manual_seed(100 + rank)
        hm = torch.rand(rows, k, 64, 48, device=dev, generator=gen)
        fl = torch.rand(rows, k, 64, 48, device=dev, generator=gen)
        center = torch.rand(rows, 2, device=dev, generator=gen) * 400
        scale = torch.rand(rows, 2, device=dev, generator=gen) * 2.8 + 0.2
        score = torch.rand(rows, device=dev, generator=gen)
        dec = mp.create_decoder("topdown_heatmap", dark_udp_refine=True)
        p = dec._params(k, 64, 48, flip_index=synth.flip_index(), shift_heatmap=False)
        for _ in range(4):
            preds, boxes = codec.topdown_decode(hm, center, scale, score, flipped=fl, params=p,
                                                gather=g)
            got_p, got_b = g.last_ticket().wait()
            plain_p, plain_b = codec.topdown_decode(hm, center, scale, score, flipped=fl, params=p)
            want_p, want_b = pdist.all_gather_keypoints(plain_p, plain_b, world * rows)
            torch.cuda.synchronize()
            assert torch.equal(preds, plain_p) and torch.equal(boxes, plain_b)
            assert torch.equal(got_p, want_p) and torch.equal(got_b, want_b)
        np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([int(g.multicast)]))
    finally:
        dist.destroy_process_group()


def test_peer_gather_matches_nccl_all_gather(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    world, rows, k = 2, 257, 17
    mp.spawn(_worker, args=(world, _free_port(), rows, k, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert os.path.exists(tmp_path / f"ok{r}.npy")
