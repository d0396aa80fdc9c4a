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


class PendingGather:
    """Handle of an all-gather in flight (``all_gather_keypoints(..., async_op=True)``): the
    collective runs on the backend's own stream while the caller's stream goes on with the
    next batch; ``wait()`` orders the caller's stream after it and returns the results."""

    def __init__(self, work, out, sizes, longest, num_joints):
        self._work, self._out, self._sizes = work, out, sizes
        self._longest, self._k = longest, num_joints

    def wait(self):
        if self._work is not None:
            self._work.wait()
            self._work = None
        sizes, longest, out = self._sizes, self._longest, self._out
        if all(s == longest for s in sizes):
            packed = out
        else:
            packed = torch.cat([out[r * longest: r * longest + sizes[r]]
                                for r in range(len(sizes))], dim=0)
        return unpack_results(packed, self._k)


def all_gather_keypoints(preds: torch.Tensor, boxes: torch.Tensor, total: int,
                         group: Optional[dist.ProcessGroup] = None, async_op: bool = False):
    """Every rank passes the results of its `shard_range` block; every rank gets
    (all_preds [total,K,3], all_boxes [total,6]) in global crop order -- or, with
    ``async_op=True``, a `PendingGather` whose ``wait()`` returns them."""
    if not dist.is_available() or not dist.is_initialized():
        if async_op:
            n = preds.shape[0]
            return PendingGather(None, pack_results(preds, boxes), [n], n, preds.shape[1])
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
    work = None
    try:
        work = dist.all_gather_into_tensor(out, local, group=group, async_op=async_op)
    except (RuntimeError, NotImplementedError):  # backend without the flat variant
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local, group=group)
        out = torch.cat(parts, dim=0)
        work = None
    pending = PendingGather(work if async_op else None, out, sizes, longest, k)
    return pending if async_op else pending.wait()


class GatherTicket:
    """One gather in flight (``PeerGather.gather_async``): this rank's rows are on their way
    into every rank's table; ``wait()`` orders the caller's current stream after the arrival
    of EVERY rank's rows of that step and returns the views of the gathered table."""

    def __init__(self, owner, step: int, table: torch.Tensor):
        self._owner, self.step, self._table = owner, step, table

    def wait(self):
        self._owner._wait(self.step)
        return unpack_results(self._table, self._owner.k)


class PeerGather:
    """The keypoint all-gather as ONE kernel of direct stores into peer memory
    (``pc_scatter_results_signal``) instead of an NCCL collective: every rank's gathered table
    is a symmetric-memory allocation, and each rank writes its block of rows into all of
    them -- through the NVSwitch multicast mapping when the allocation has one, else through
    the peer-mapped addresses over NVLink.  The kernel's last CTA then publishes the step
    number in a flag word on every rank; a rank waits for the flags of a step
    (``pc_wait_peer_flags``, a one-warp kernel) only when it reads that step's table.

    ``gather()`` scatters and waits at once.  ``gather_async()`` returns a ``GatherTicket``;
    calling ``ticket.wait()`` one step later takes the wait for the slowest rank off the
    critical path of the step (bench.py does that).  THREE tables alternate: a peer may
    overwrite the table of step s as soon as it has seen this rank's flag of step s + 2, which
    this rank only publishes after everything it enqueued before ``gather_async`` of step
    s + 2 -- so the views a ticket returns are valid until the second ``gather*`` call after
    the one that made the ticket, for work enqueued on the same stream.
    """

    TABLES = 3

    def __init__(self, rows_per_rank: int, num_joints: int, device: torch.device,
                 group: Optional[dist.ProcessGroup] = None):
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib

        self._lib, self._ctypes = _lib, ctypes
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.rows, self.k = int(rows_per_rank), int(num_joints)
        self.width = self.k * 3 + 6
        grp = group if group is not None else dist.group.WORLD
        self.tables, self.handles, self._peers, self._mc = [], [], [], []
        for _ in range(self.TABLES):
            t = symm_mem.empty((self.world * self.rows, self.width), dtype=torch.float32,
                               device=device)
            h = symm_mem.rendezvous(t, group=grp)
            self.tables.append(t)
            self.handles.append(h)
            ptrs = [int(p) for p in h.buffer_ptrs]
            self._peers.append((ctypes.c_void_p * self.world)(*ptrs))
            mc = int(getattr(h, "multicast_ptr", 0) or 0)
            self._mc.append(mc)
        self.multicast = all(m != 0 for m in self._mc)
        # flags[r] on this rank = the last step whose rows rank r has stored here
        self._flags = symm_mem.empty((max(self.world, 32),), dtype=torch.int32, device=device)
        self._flags.zero_()
        fh = symm_mem.rendezvous(self._flags, group=grp)
        self._flag_handle = fh
        self._peer_flags = (ctypes.c_void_p * self.world)(*[int(p) for p in fh.buffer_ptrs])
        self._counter = torch.zeros(1, dtype=torch.int32, device=device)
        # the step number the kernels publish lives in device memory (a captured CUDA graph of
        # a step then signals a new number at every replay); `_step` mirrors it on the host
        self._dstep = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        fh.barrier(channel=0)            # nobody signals before everybody has zeroed its flags
        self._step = 0                   # steps count from 1 (flags start at 0)
        self._waited = 0

    def gather_async(self, preds: torch.Tensor, boxes: torch.Tensor) -> GatherTicket:
        """preds f32 [rows,K,3], boxes f32 [rows,6] of this rank: one kernel on the current
        stream stores them into every rank's table and signals; nothing waits."""
        if preds.shape[0] != self.rows or boxes.shape[0] != self.rows:
            raise ValueError(f"expected {self.rows} rows per rank, got {preds.shape[0]}")
        if not (preds.is_cuda and preds.is_contiguous() and boxes.is_contiguous()):
            raise ValueError("preds / boxes must be contiguous CUDA tensors")
        self._step += 1
        i = self._step % self.TABLES
        lib = self._lib
        lib.call("pc_scatter_results_signal", lib.device_ptr(preds), lib.device_ptr(boxes),
                 self._peers[i], self.world, self._mc[i] if self.multicast else None,
                 self.rank * self.rows, self.k, self.rows, self._peer_flags, self.world,
                 self.rank, lib.device_ptr(self._dstep), lib.device_ptr(self._counter),
                 lib.current_stream())
        return GatherTicket(self, self._step, self.tables[i])

    def next_target(self, rows: int, num_joints: int):
        """The ``pc_gather_target`` of the NEXT step, for a producer kernel that stores its
        results into the gathered tables itself (``codec.topdown_decode(..., gather=self)``:
        decode and all-gather as one kernel).  Counts as this step's scatter."""
        if rows != self.rows or num_joints != self.k:
            raise ValueError(f"this gather holds {self.rows} rows of {self.k} joints per rank")
        self._step += 1
        i = self._step % self.TABLES
        lib = self._lib
        t = lib.GatherTarget()
        t.h_peer_tables = self._peers[i]
        t.num_peers = self.world
        t.d_multicast_table = self._mc[i] if self.multicast else None
        t.row_offset = self.rank * self.rows
        t.h_peer_flags = self._peer_flags
        t.num_flag_peers = self.world
        t.my_rank = self.rank
        t.d_step = lib.device_ptr(self._dstep)
        t.d_counter = lib.device_ptr(self._counter)
        return t

    def last_ticket(self) -> GatherTicket:
        """Ticket of the latest scatter (made by ``gather_async`` or by a fused producer)."""
        return GatherTicket(self, self._step, self.tables[self._step % self.TABLES])

    def _wait(self, step: int) -> None:
        if step <= self._waited:         # flags only grow: a later wait covers earlier steps
            return
        lib = self._lib
        lib.call("pc_wait_peer_flags", lib.device_ptr(self._flags), self.world,
                 lib.device_ptr(self._dstep), self._step - step, lib.current_stream())
        self._waited = step

    def wait_lag(self, lag: int = 0) -> None:
        """Order the current stream after the arrival of every rank's rows of the scatter
        issued `lag` scatters ago (0: the latest).  Unlike a ticket this needs no host state
        per step, so a captured CUDA graph can hold it: "scatter; wait_lag(1)" per step."""
        if lag < 0:
            raise ValueError("lag must be >= 0")
        lib = self._lib
        lib.call("pc_wait_peer_flags", lib.device_ptr(self._flags), self.world,
                 lib.device_ptr(self._dstep), int(lag), lib.current_stream())
        self._waited = max(self._waited, self._step - lag)

    def table_of_lag(self, lag: int = 0):
        """Views (all_preds, all_boxes) of the table the scatter `lag` scatters ago wrote."""
        return unpack_results(self.tables[(self._step - lag) % self.TABLES], self.k)

    def captured(self, steps: int) -> None:
        """`steps` gather steps were just CAPTURED into a CUDA graph (nothing ran on the
        device): take them back from the host mirror.  Call `replayed(steps)` per replay."""
        if steps % self.TABLES:
            raise ValueError(f"a captured graph must hold a multiple of {self.TABLES} gather steps")
        self._step -= steps
        self._waited = min(self._waited, self._step)

    def replayed(self, steps: int) -> None:
        """A captured CUDA graph holding `steps` gather steps (a multiple of TABLES, so that
        the table rotation closes) was replayed once more: advance the host mirror of the
        device's step number.  Tickets made during the capture are not valid afterwards."""
        if steps % self.TABLES:
            raise ValueError(f"a captured graph must hold a multiple of {self.TABLES} gather steps")
        self._step += steps
        self._waited += steps

    def gather(self, preds: torch.Tensor, boxes: torch.Tensor):
        """-> (all_preds [world*rows,K,3], all_boxes [world*rows,6]) once every rank's rows
        have landed (scatter + wait on the current stream)."""
        return self.gather_async(preds, boxes).wait()


def make_gatherer(rows_per_rank: int, num_joints: int, device: torch.device,
                  group: Optional[dist.ProcessGroup] = None):
    """``PeerGather`` when symmetric memory works on this box, else None (callers fall back
    to ``all_gather_keypoints``, the NCCL collective)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2:
        return None
    try:
        return PeerGather(rows_per_rank, num_joints, device, group)
    except Exception as e:  # no P2P / VMM support, old driver, ...
        import warnings

        warnings.warn(f"peer-memory gather unavailable ({type(e).__name__}: {e}); using NCCL")
        return None
