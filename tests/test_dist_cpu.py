"""World-size-2 (and 3) gloo tests of the sharding + all-gather host logic (CPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mindpose_b200 import dist as pdist


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 4096, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [pdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        pdist.shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = pdist.shard_range(total, rank, world)
        idx = torch.arange(lo, hi, dtype=torch.float32)
        # results that encode the global crop index, so order is checkable
        preds = idx[:, None, None] * 10 + torch.arange(k * 3, dtype=torch.float32).reshape(1, k, 3)
        boxes = idx[:, None] * 100 + torch.arange(6, dtype=torch.float32)[None]
        if os.environ.get("PC_TEST_ASYNC_GATHER") == "1":
            handles = [pdist.all_gather_keypoints(preds + j, boxes, total, async_op=True)
                       for j in range(3)]          # several gathers in flight
            outs = [h.wait() for h in handles]
            for j, (p_j, _) in enumerate(outs):
                assert torch.equal(p_j - j, outs[0][0])
            all_p, all_b = outs[0]
        else:
            all_p, all_b = pdist.all_gather_keypoints(preds, boxes, total)
        np.save(os.path.join(out_dir, f"p{rank}.npy"), all_p.numpy())
        np.save(os.path.join(out_dir, f"b{rank}.npy"), all_b.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total,async_op", [(2, 64, False), (2, 33, False), (3, 10, False),
                                                   (2, 33, True)])
def test_all_gather_keypoints_gloo(tmp_path, monkeypatch, world, total, async_op):
    k = 17
    monkeypatch.setenv("PC_TEST_ASYNC_GATHER", "1" if async_op else "0")
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total, k, str(tmp_path)), nprocs=world, join=True)
    idx = np.arange(total, dtype=np.float32)
    want_p = idx[:, None, None] * 10 + np.arange(k * 3, dtype=np.float32).reshape(1, k, 3)
    want_b = idx[:, None] * 100 + np.arange(6, dtype=np.float32)[None]
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"p{r}.npy"), want_p)
        assert np.array_equal(np.load(tmp_path / f"b{r}.npy"), want_b)


def test_gather_without_process_group_is_identity():
    p, b = torch.zeros(4, 17, 3), torch.zeros(4, 6)
    q, c = pdist.all_gather_keypoints(p, b, 4)
    assert q is p and c is b
