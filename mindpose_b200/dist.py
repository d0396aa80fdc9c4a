"""Multi-GPU plumbing of the codec: one process per GPU, work sharded by crop /
image, and ONE collective -- an all-gather of the decoded keypoints so that every
rank (or rank 0) can run COCO-style evaluation.

The reference shards its training dataset the same way
(``GeneratorDataset(num_shards=device_num, shard_id=rank_id)``,
mindpose/data/data_factory.py:59-66) and evaluates on rank 0 only
(mindpose/callbacks/eval_callback.py:142-145); the gather is the one new step.
Encode / warp / decode need no communication at all: every crop is independent.
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of global indices owned by `rank`: sizes differ by at most 1,
    the first ``n % world`` ranks hold the longer blocks."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n: int, world: int):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def pack_results(preds: torch.Tensor, boxes: torch.Tensor) -> torch.Tensor:
    """[n,K,3] + [n,6] -> [n, K*3 + 6] (228 bytes per crop for K = 17)."""
    n = preds.shape[0]
    return torch.cat([preds.reshape(n, -1), boxes.reshape(n, -1)], dim=1).contiguous()


def unpack_results(packed: torch.Tensor, num_joints: int):
    n = packed.shape[0]
    return (packed[:, : num_joints * 3].reshape(n, num_joints, 3),
            packed[:, num_joints * 3:].reshape(n, 6))


def all_gather_keypoints(preds: torch.Tensor, boxes: torch.Tensor, total: int,
                         group: Optional[dist.ProcessGroup] = None):
    """Every rank passes the results of its `shard_range` block; every rank gets
    (all_preds [total,K,3], all_boxes [total,6]) in global crop order."""
    if not dist.is_available() or not dist.is_initialized():
        return preds, boxes
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    k = preds.shape[1]
    sizes = shard_sizes(total, world)
    if preds.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {preds.shape[0]} crops, its shard has {sizes[rank]}")
    local = pack_results(preds, boxes)
    width = local.shape[1]
    longest = max(sizes)
    if local.shape[0] < longest:  # ragged tail: pad to the longest shard, trim after
        pad = torch.zeros((longest - local.shape[0], width), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    out = torch.empty((world * longest, width), dtype=local.dtype, device=local.device)
    try:
        dist.all_gather_into_tensor(out, local, group=group)
    except (RuntimeError, NotImplementedError):  # backend without the flat variant
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local, group=group)
        out = torch.cat(parts, dim=0)
    if all(s == longest for s in sizes):
        packed = out
    else:
        packed = torch.cat([out[r * longest: r * longest + sizes[r]] for r in range(world)], dim=0)
    return unpack_results(packed, k)
