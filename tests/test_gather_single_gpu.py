"""The peer-gather kernels on ONE GPU, through the C ABI with this rank as its only peer (plain
device buffers stand in for the symmetric-memory tables): pc_scatter_results_signal,
pc_wait_peer_flags and pc_topdown_decode_gather run on the 1-GPU test box too.  The multi-rank
behaviour (multicast, deferred waits, three tables) is tests/test_peer_gather_gpu.py (2 GPUs)
and bench.py's `gather_verified` at every N."""
import ctypes

import numpy as np
import pytest
import torch

import mindpose_b200 as mp
from mindpose_b200 import _lib, codec, synth

pytestmark = pytest.mark.gpu


def _self_peer(dev, rows, k, world=3, rank=1):
    """A gathered table of `world` ranks' worth of rows of which only rank `rank` (this GPU)
    exists: its rows land at row_offset = rank * rows; everything else must stay untouched."""
    width = k * 3 + 6
    table = torch.full((world * rows, width), -7.0, device=dev)
    flags = torch.zeros(32, dtype=torch.int32, device=dev)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    counter = torch.zeros(1, dtype=torch.int32, device=dev)
    peers = (ctypes.c_void_p * 1)(table.data_ptr())
    # flag arrays: one "rank" (this one) publishes into word `rank` of its own array
    pflags = (ctypes.c_void_p * (rank + 1))(*([flags.data_ptr()] * (rank + 1)))
    return table, flags, step, counter, peers, pflags


def test_scatter_signal_and_wait_with_self_as_peer(cuda_device):
    dev = cuda_device
    rows, k, world, rank = 257, 17, 3, 1
    table, flags, step, counter, peers, pflags = _self_peer(dev, rows, k, world, rank)
    g = torch.Generator(device=dev).manual_seed(0)
    for it in range(1, 4):
        preds = torch.rand(rows, k, 3, device=dev, generator=g)
        boxes = torch.rand(rows, 6, device=dev, generator=g)
        _lib.call("pc_scatter_results_signal", _lib.device_ptr(preds), _lib.device_ptr(boxes), peers,
                  1, None, rank * rows, k, rows, pflags, rank + 1, rank, _lib.device_ptr(step),
                  _lib.device_ptr(counter), _lib.current_stream())
        # only word `rank` is awaited: num_peers = rank + 1 would also wait for the ranks that
        # do not exist here, so wait on the published word through a one-word view
        _lib.call("pc_wait_peer_flags", flags.data_ptr() + 4 * rank, 1, _lib.device_ptr(step), 0,
                  _lib.current_stream())
        torch.cuda.synchronize()
        assert int(step.item()) == it and int(flags[rank].item()) == it and int(counter.item()) == 0
        mine = table[rank * rows:(rank + 1) * rows]
        assert torch.equal(mine[:, :k * 3].reshape(rows, k, 3), preds)
        assert torch.equal(mine[:, k * 3:], boxes)
        assert bool((table[:rank * rows] == -7).all()) and bool((table[(rank + 1) * rows:] == -7).all())
    # an empty shard still signals
    _lib.call("pc_scatter_results_signal", None, None, peers, 1, None, 0, k, 0, pflags, rank + 1, rank,
              _lib.device_ptr(step), _lib.device_ptr(counter), _lib.current_stream())
    torch.cuda.synchronize()
    assert int(flags[rank].item()) == 4
    # rows not a multiple of 4 / odd offsets take the scalar store path
    table.fill_(-7.0)
    preds = torch.rand(rows, k, 3, device=dev, generator=g)
    boxes = torch.rand(rows, 6, device=dev, generator=g)
    _lib.call("pc_scatter_results", _lib.device_ptr(preds), _lib.device_ptr(boxes), peers, 1, None,
              3, k, rows, _lib.current_stream())
    torch.cuda.synchronize()
    assert torch.equal(table[3:3 + rows, :k * 3].reshape(rows, k, 3), preds)
    assert torch.equal(table[3:3 + rows, k * 3:], boxes)


@pytest.mark.parametrize("flip,dark", [(True, True), (False, False)])
def test_decode_gather_as_one_kernel_with_self_as_peer(cuda_device, flip, dark):
    """pc_topdown_decode_gather: the local results equal pc_topdown_decode's, and the same
    values sit in this rank's rows of the gathered table when the step flag is published."""
    dev = cuda_device
    rows, k, h, w, world, rank = 300, 17, 64, 48, 2, 1
    table, flags, step, counter, peers, pflags = _self_peer(dev, rows, k, world, rank)
    maps, _ = synth.blob_heatmaps(rows, k, h, w, seed=9)
    flipped = synth.flipped_pair(maps, seed=9)
    center, scale, score = synth.crop_geometry(rows, seed=9)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    hm, fl, c, s, sc = t(maps), t(flipped), t(center), t(scale), t(score)
    dec = mp.create_decoder("topdown_heatmap", dark_udp_refine=dark)
    p = dec._params(k, h, w, flip_index=synth.flip_index() if flip else None, shift_heatmap=False)
    want_p, want_b = codec.topdown_decode(hm, c, s, sc, flipped=fl if flip else None, params=p)
    target = _lib.GatherTarget()
    target.h_peer_tables = peers
    target.num_peers = 1
    target.d_multicast_table = None
    target.row_offset = rank * rows
    target.h_peer_flags = pflags
    target.num_flag_peers = rank + 1
    target.my_rank = rank
    target.d_step = _lib.device_ptr(step)
    target.d_counter = _lib.device_ptr(counter)
    got_p = torch.empty_like(want_p)
    got_b = torch.empty_like(want_b)
    for it in (1, 2):
        _lib.call("pc_topdown_decode_gather", _lib.device_ptr(hm), _lib.device_ptr(fl) if flip else None,
                  _lib.device_ptr(c), _lib.device_ptr(s), _lib.device_ptr(sc), _lib.device_ptr(got_p),
                  _lib.device_ptr(got_b), ctypes.byref(p), rows, ctypes.byref(target),
                  _lib.current_stream())
        _lib.call("pc_wait_peer_flags", flags.data_ptr() + 4 * rank, 1, _lib.device_ptr(step), 0,
                  _lib.current_stream())
        torch.cuda.synchronize()
        assert int(flags[rank].item()) == it and int(counter.item()) == 0
        assert torch.equal(got_p, want_p) and torch.equal(got_b, want_b)
        mine = table[rank * rows:(rank + 1) * rows]
        assert torch.equal(mine[:, :k * 3].reshape(rows, k, 3), want_p)
        assert torch.equal(mine[:, k * 3:], want_b)
        assert bool((table[:rank * rows] == -7).all())
    with pytest.raises(ValueError):
        target.num_flag_peers = 0
        _lib.call("pc_topdown_decode_gather", _lib.device_ptr(hm), None, _lib.device_ptr(c),
                  _lib.device_ptr(s), _lib.device_ptr(sc), _lib.device_ptr(got_p),
                  _lib.device_ptr(got_b), ctypes.byref(p), rows, ctypes.byref(target),
                  _lib.current_stream())
