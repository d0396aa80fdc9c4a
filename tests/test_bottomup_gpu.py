"""Parity of the CUDA bottom-up codec (decode, grouping, back-projection) against the
oracle (B200 only).  Bars: top-k values / indices and grouping assignments bit-exact."""
import numpy as np
import pytest
import torch

import mindpose_b200 as mp
from mindpose_b200 import bottomup, synth
from oracle import bottomup_decode as bd
from oracle import grouping
from oracle.gen_golden_bottomup import grouping_inputs

pytestmark = pytest.mark.gpu


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _split_counts(flat, counts, k=17, width=4):
    out, off = [], 0
    for c in counts:
        size = int(c) * k * width
        out.append(flat[off:off + size].reshape(int(c), k, width))
        off += size
    return out


@pytest.mark.parametrize("h0,mask_hw,use_nms,nms_kernel", [
    (32, (128, 128), True, 3), (32, (128, 128), False, 5), (64, (256, 256), True, 5),
    (48, (100, 210), True, 3),
    (160, (640, 640), True, 3),     # W = 320: 16 columns per lane
    (160, (200, 300), False, 3),
    (20, (80, 80), True, 3),        # H = 40 < 3 rows per warp band, W = 40 (generic: 40 % 4 == 0, C = 4)
    (128, (512, 512), True, 1)])    # BASELINE config 4 shape, 1x1 pool
def test_decode_matches_oracle(cuda_device, h0, mask_hw, use_nms, nms_kernel):
    n = 3
    d = synth.bottomup_outputs(n, 17, h0, h0, mask_hw=mask_hw, seed=h0, max_people=5)
    want = bd.decode([d["out0"], d["out1"]], d["mask"], use_nms=use_nms, nms_kernel=nms_kernel,
                     max_num=30)
    dec = mp.create_decoder("bottomup_heatmap_ae", use_nms=use_nms, nms_kernel=nms_kernel,
                            max_num=30)
    got = dec([_t(d["out0"], cuda_device), _t(d["out1"], cuda_device)], _t(d["mask"], cuda_device))
    names = ["val_k", "tag_k", "ind_k", "heatmap_raw", "tagging_heatmap"]
    assert len(got) == 5
    for name, g, w in zip(names, got, want):
        assert g.shape == w.shape, name
        assert np.array_equal(g.cpu().numpy(), w), name


def test_decode_single_stage_and_public_decode(cuda_device):
    n, k, h = 2, 17, 64
    rng = np.random.RandomState(0)
    out0 = rng.uniform(-0.2, 1, (n, 2 * k, h, h)).astype(np.float32)
    mask = (rng.random_sample((n, 128, 128)) > 0.1).astype(np.uint8)
    want = bd.decode([out0], mask, num_stages=1, with_ae_loss=(True,), use_nms=True,
                     nms_kernel=3, max_num=20)
    dec = mp.create_decoder("bottomup_heatmap_ae", num_stages=1, with_ae_loss=[True],
                            use_nms=True, nms_kernel=3, max_num=20)
    t0 = _t(out0, cuda_device)
    heat, tag = dec.decouple_output([t0])
    assert heat[0].shape == (n, k, h, h) and tag[0].shape == (n, k, h, h)
    got = dec.decode(heat, tag, _t(mask, cuda_device))
    for g, w in zip(got[:3], want[:3]):
        assert np.array_equal(g.cpu().numpy(), w)


def test_decode_accepts_batch_sliced_views(cuda_device):
    """`decode()` on views cut out of a larger batch (big[2:4]): the contiguous base of such
    a view holds OTHER images, so it must never be read in place of the view."""
    n, k, h0 = 6, 17, 32
    d = synth.bottomup_outputs(n, k, h0, h0, mask_hw=(128, 128), seed=11, max_people=4)
    dec = mp.create_decoder("bottomup_heatmap_ae", use_nms=True, nms_kernel=3, max_num=30)
    big0, big1 = _t(d["out0"], cuda_device), _t(d["out1"], cuda_device)
    mask = _t(d["mask"], cuda_device)
    sl = slice(2, 4)
    want = bd.decode([d["out0"][sl], d["out1"][sl]], d["mask"][sl], use_nms=True, nms_kernel=3,
                     max_num=30)
    heat, tag = dec.decouple_output([big0[sl], big1[sl]])
    got = dec.decode(heat, tag, mask[sl])
    for g, w in zip(got, want):
        assert np.array_equal(g.cpu().numpy(), w)
    # a channel-sliced second stage (not contiguous) is packed, not misread
    wide = torch.cat([big1, big1.flip(1)], dim=1)
    heat2 = [heat[0], wide[sl, :k]]
    got = dec.decode(heat2, tag, mask[sl])
    for g, w in zip(got, want):
        assert np.array_equal(g.cpu().numpy(), w)
    with pytest.raises(ValueError):
        dec.decode([heat[0], big1[0:3]], tag, mask[sl])        # image counts differ


@pytest.mark.parametrize("h0,stages", [(32, 2), (128, 2), (20, 2), (64, 1)])
def test_decode_shared_tag_plane(cuda_device, h0, stages):
    """tag_per_joint=False: the first-stage output has K + 1 channels and every joint gathers
    its tags from the one shared plane (the reference's broadcast, bottom_up_decoder.py:
    159-160); all three decode kernels (generic, fast, pair scan) by shape."""
    n, k = 3, 17
    d = synth.bottomup_outputs(n, k, h0, h0, mask_hw=(4 * h0, 4 * h0), seed=h0 + 1, max_people=5)
    out0 = np.ascontiguousarray(d["out0"][:, :k + 1])          # heat | ONE tag plane
    if stages == 1:
        outs = [out0]
        kw = dict(num_stages=1, with_ae_loss=(True,))
    else:
        outs = [out0, d["out1"]]
        kw = dict(num_stages=2, with_ae_loss=(True, False))
    want = bd.decode(outs, d["mask"], use_nms=True, nms_kernel=3, max_num=30,
                     tag_per_joint=False, **kw)
    dec = mp.create_decoder("bottomup_heatmap_ae", use_nms=True, nms_kernel=3, max_num=30,
                            tag_per_joint=False, num_stages=kw["num_stages"],
                            with_ae_loss=list(kw["with_ae_loss"]))
    got = dec([_t(o, cuda_device) for o in outs], _t(d["mask"], cuda_device))
    assert got[4].shape == (n, 1, got[3].shape[2], got[3].shape[3], 1)
    for name, g, w in zip(["val_k", "tag_k", "ind_k", "heatmap_raw", "tagging_heatmap"], got, want):
        assert g.shape == w.shape, name
        assert np.array_equal(g.cpu().numpy(), w), name


def test_decode_ties_and_flat_maps(cuda_device):
    """Constant / all-zero planes: every pixel survives NMS, top-k = lowest indices."""
    n, k, h0 = 1, 17, 16
    out0 = np.zeros((n, 2 * k, h0, h0), np.float32)
    out1 = np.zeros((n, k, 2 * h0, 2 * h0), np.float32)
    out1[0, 1] = 0.5
    out1[0, 2, 5, 7] = out1[0, 2, 20, 3] = 0.9
    out1[0, 3] = -0.3                        # negative plateau: zeros outrank it nowhere
    out0[0, k:] = np.arange(h0 * h0, dtype=np.float32).reshape(h0, h0)
    mask = np.ones((n, 4 * h0, 4 * h0), np.uint8)
    mask[0, :8, :] = 0
    want = bd.decode([out0, out1], mask, use_nms=True, nms_kernel=3, max_num=30)
    dec = mp.create_decoder("bottomup_heatmap_ae", use_nms=True, nms_kernel=3, max_num=30)
    got = dec([_t(out0, cuda_device), _t(out1, cuda_device)], _t(mask, cuda_device))
    for g, w in zip(got[:3], want[:3]):
        assert np.array_equal(g.cpu().numpy(), w)


@pytest.mark.parametrize("h0,w0,mask_hw,use_nms,nms_kernel", [
    (96, 96, (384, 384), True, 3),      # W = 192: 24 active lanes
    (72, 72, (288, 288), True, 3),      # W = 144, bands of 18 rows
    (100, 100, (400, 400), True, 3),    # H = 200: ragged last band
    (30, 80, (120, 320), True, 3),      # non-square, bands of 8 rows
    (30, 80, (90, 320), True, 3),       # mask rows at 1.5x
    (30, 80, (120, 320), False, 3),
    (128, 128, (512, 512), True, 3)])
def test_decode_pair_scan_matches_oracle(cuda_device, h0, w0, mask_hw, use_nms, nms_kernel):
    """Shapes that take bottomup_decode_pairs_kernel (two stages at half size, 2x mask width,
    128 < W <= 256): all five outputs bit-exact, masks with several zero rectangles."""
    n = 3
    d = synth.bottomup_outputs(n, 17, h0, w0, mask_hw=mask_hw, seed=h0 + w0, max_people=6)
    rng = np.random.RandomState(h0)
    for i in range(n):
        for _ in range(4):
            y0, x0 = rng.randint(0, mask_hw[0] - 8), rng.randint(0, mask_hw[1] - 8)
            d["mask"][i, y0:y0 + rng.randint(1, 40), x0:x0 + rng.randint(1, 90)] = 0
    d["mask"][1, :, :5] = 0
    d["mask"][2, -3:, :] = 0
    want = bd.decode([d["out0"], d["out1"]], d["mask"], use_nms=use_nms, nms_kernel=nms_kernel,
                     max_num=30)
    dec = mp.create_decoder("bottomup_heatmap_ae", use_nms=use_nms, nms_kernel=nms_kernel,
                            max_num=30)
    got = dec([_t(d["out0"], cuda_device), _t(d["out1"], cuda_device)], _t(d["mask"], cuda_device))
    for name, g, w in zip(["val_k", "tag_k", "ind_k", "heatmap_raw", "tagging_heatmap"], got, want):
        assert np.array_equal(g.cpu().numpy(), w), name


@pytest.mark.parametrize("max_num", [30, 7])
def test_decode_pair_scan_sparse_and_tied_planes(cuda_device, max_num):
    """Planes with fewer than M positive survivors, plateaus, exact ties and one lane region
    holding many of the top M: the exact second pass must take over."""
    n, k, h0, w0 = 2, 17, 40, 80
    rng = np.random.RandomState(5)
    out0 = np.zeros((n, 2 * k, h0, w0), np.float32)
    out1 = np.zeros((n, k, 2 * h0, 2 * w0), np.float32)
    out0[:, k:] = rng.uniform(-1, 1, (n, k, h0, w0))
    out1[0, 1] = 0.5                                   # plateau: every pixel survives
    out1[0, 2, 5, 7] = out1[0, 2, 60, 130] = 0.9       # two spikes, zeros fill the rest
    out1[0, 3] = -0.3                                  # negative plateau
    out1[0, 4] = rng.uniform(-0.5, -0.1, (2 * h0, 2 * w0))   # all negative: zeros outrank survivors
    for j in range(12):                                # 12 of the top M in one 8 x 10 lane region
        out1[0, 5, 20 + 2 * (j // 4), 40 + 2 * (j % 4)] = 0.5 + 0.01 * j
    out1[0, 5] += rng.uniform(0, 0.01, (2 * h0, 2 * w0)).astype(np.float32)
    out1[0, 6] = np.round(rng.uniform(0, 4, (2 * h0, 2 * w0))) / 4   # heavy ties
    out0[0, 6] = np.round(rng.uniform(0, 4, (h0, w0))) / 4
    out1[1] = rng.uniform(-0.02, 0.02, (k, 2 * h0, 2 * w0))
    out0[1, :k] = rng.uniform(-0.02, 0.02, (k, h0, w0))
    mask = np.ones((n, 4 * h0, 4 * w0), np.uint8)
    mask[0, :9, :] = 0
    mask[1, 50:70, 100:300] = 0
    want = bd.decode([out0, out1], mask, use_nms=True, nms_kernel=3, max_num=max_num)
    dec = mp.create_decoder("bottomup_heatmap_ae", use_nms=True, nms_kernel=3, max_num=max_num)
    got = dec([_t(out0, cuda_device), _t(out1, cuda_device)], _t(mask, cuda_device))
    for name, g, w in zip(["val_k", "tag_k", "ind_k", "heatmap_raw"], got, want):
        assert np.array_equal(g.cpu().numpy(), w), name


def test_decode_full_size_properties(cuda_device):
    """BASELINE config 4 shapes (64 x [34,128,128] + [17,256,256], 512^2 mask): values sorted,
    indices in range and consistent with heatmap_raw; spot-check 2 images against the oracle."""
    n = 64
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(0)
    out0 = torch.rand(n, 34, 128, 128, device=dev, generator=g) * 0.02
    out1 = torch.rand(n, 17, 256, 256, device=dev, generator=g) * 0.02
    ys = torch.randint(8, 248, (n, 17, 6), device=dev, generator=g)
    xs = torch.randint(8, 248, (n, 17, 6), device=dev, generator=g)
    for j in range(6):
        out1[torch.arange(n)[:, None], torch.arange(17)[None, :], ys[..., j], xs[..., j]] += 0.5 + 0.05 * j
    mask = torch.ones(n, 512, 512, dtype=torch.uint8, device=dev)
    mask[:, 100:140, 200:260] = 0
    dec = mp.create_decoder("bottomup_heatmap_ae", use_nms=True, nms_kernel=3, max_num=30)
    val_k, tag_k, ind_k, raw, tagging = dec([out0, out1], mask)
    assert torch.all(val_k[..., :-1] >= val_k[..., 1:])
    x, y = ind_k[..., 0].long(), ind_k[..., 1].long()
    assert x.min() >= 0 and x.max() < 256 and y.min() >= 0 and y.max() < 256
    picked = raw.reshape(n, 17, -1).gather(2, y * 256 + x)
    assert torch.equal(picked, val_k)          # survivors keep their raw value
    assert torch.equal(tagging.reshape(n, 17, -1).gather(2, y * 256 + x), tag_k[..., 0])
    sel = [0, 63]
    want = bd.decode([out0[sel].cpu().numpy(), out1[sel].cpu().numpy()], mask[sel].cpu().numpy(),
                     use_nms=True, nms_kernel=3, max_num=30)
    for gt, w in zip((val_k, tag_k, ind_k), want[:3]):
        assert np.array_equal(gt[sel].cpu().numpy(), w)


@pytest.mark.parametrize("h0,w0,mask_hw,stages", [
    (32, 32, (128, 128), 2),       # C = 4 fast kernel
    (128, 128, (512, 512), 2),     # pair kernel
    (20, 21, (80, 84), 2),         # generic kernel (W = 42)
    (64, 64, (128, 128), 1)])
def test_decode_shift_coordinate_matches_oracle(cuda_device, h0, w0, mask_hw, stages):
    """A17 with the reference's own pairing of offsets and candidates (oracle:
    shift_coordinate_quirk)."""
    n = 2
    if stages == 2:
        d = synth.bottomup_outputs(n, 17, h0, w0, mask_hw=mask_hw, seed=7 + h0, max_people=5)
        outs = [d["out0"], d["out1"]]
        mask = d["mask"]
        kw = {}
        okw = {}
    else:
        rng = np.random.RandomState(1)
        outs = [rng.uniform(-0.2, 1, (n, 34, h0, w0)).astype(np.float32)]
        mask = (rng.random_sample((n,) + mask_hw) > 0.1).astype(np.uint8)
        kw = dict(num_stages=1, with_ae_loss=[True])
        okw = dict(num_stages=1, with_ae_loss=(True,))
    want = bd.decode(outs, mask, use_nms=True, nms_kernel=3, max_num=30, shift_coordinate=True,
                     **okw)
    dec = mp.create_decoder("bottomup_heatmap_ae", use_nms=True, nms_kernel=3, max_num=30,
                            shift_coordinate=True, **kw)
    got = dec([_t(o, cuda_device) for o in outs], _t(mask, cuda_device))
    for name, g, w in zip(["val_k", "tag_k", "ind_k"], got, want):
        assert np.array_equal(g.cpu().numpy(), w), name
    plain = bd.decode(outs, mask, use_nms=True, nms_kernel=3, max_num=30, **okw)
    assert not np.array_equal(plain[2], want[2])      # the shift did something


@pytest.mark.parametrize("max_num", [33, 50, 64])
@pytest.mark.parametrize("h0,w0,stages", [(32, 32, 2), (24, 20, 2), (64, 64, 1)])
def test_decode_and_grouping_max_num_above_32(cuda_device, max_num, h0, w0, stages):
    """max_num up to 64 (the reference's top-k takes any k, bottom_up_decoder.py:131-171):
    decode (with the coordinate shift), then grouping over 64 detections per joint."""
    n = 2
    rng = np.random.RandomState(max_num + h0)
    if stages == 2:
        d = synth.bottomup_outputs(n, 17, h0, w0, mask_hw=(4 * h0, 4 * w0), seed=max_num + h0,
                                   max_people=12)
        outs, mask = [d["out0"], d["out1"]], d["mask"]
        kw, okw = {}, {}
    else:
        outs = [rng.uniform(-0.2, 1, (n, 34, h0, w0)).astype(np.float32)]
        mask = (rng.random_sample((n, 2 * h0, 2 * w0)) > 0.1).astype(np.uint8)
        kw, okw = dict(num_stages=1, with_ae_loss=[True]), dict(num_stages=1, with_ae_loss=(True,))
    for shift in (False, True):
        want = bd.decode(outs, mask, use_nms=True, nms_kernel=3, max_num=max_num,
                         shift_coordinate=shift, **okw)
        dec = mp.create_decoder("bottomup_heatmap_ae", use_nms=True, nms_kernel=3, max_num=max_num,
                                shift_coordinate=shift, **kw)
        got = dec([_t(o, cuda_device) for o in outs], _t(mask, cuda_device))
        for name, g, w in zip(["val_k", "tag_k", "ind_k", "raw", "tagging"], got, want):
            assert np.array_equal(g.cpu().numpy(), w), (name, shift)
    val_k, tag_k, ind_k = want[:3]
    room = bottomup.max_group_capacity(17, max_num)
    ans, num, _ = bottomup.group_by_tag(_t(val_k, cuda_device), _t(tag_k, cuda_device),
                                        _t(ind_k, cuda_device), synth.COCO_JOINT_ORDER,
                                        max_groups=room)
    ans, num = ans.cpu().numpy(), num.cpu().numpy()
    for i in range(n):
        w = grouping.match_by_tag(val_k[i], tag_k[i], ind_k[i], synth.COCO_JOINT_ORDER)
        p = 0 if w.ndim == 1 else w.shape[0]
        assert num[i] == p, i
        if p:
            assert np.array_equal(ans[i, :p], w), i
    with pytest.raises(ValueError, match="max_num"):
        mp.create_decoder("bottomup_heatmap_ae", max_num=65, **kw)(
            [_t(o, cuda_device) for o in outs], _t(mask, cuda_device))


# ------------------------------------------------------------------- grouping
@pytest.mark.parametrize("mode", ["people", "ties", "crowded"])
@pytest.mark.parametrize("rounded", [True, False])
def test_grouping_matches_reference_golden(cuda_device, golden, mode, rounded):
    g = golden("match_ref.npz")
    name = f"{mode}_{'rounded' if rounded else 'exact'}"
    want = _split_counts(g[f"ans_{name}"], g[f"counts_{name}"])
    val, tag, ind = g[f"val_{mode}"], g[f"tag_{mode}"], g[f"ind_{mode}"]
    ans, num, scores = bottomup.group_by_tag(
        _t(val, cuda_device), _t(tag, cuda_device), _t(ind, cuda_device),
        synth.COCO_JOINT_ORDER, vis_thr=0.1, tag_thr=1.0, use_rounded_norm=rounded)
    ans, num, scores = ans.cpu().numpy(), num.cpu().numpy(), scores.cpu().numpy()
    for i in range(val.shape[0]):
        assert num[i] == want[i].shape[0], (mode, rounded, i)
        assert np.array_equal(ans[i, :num[i]], want[i]), (mode, rounded, i)
        assert scores[i, :num[i]].tolist() == [y[:, 2].mean() for y in want[i]]


def test_grouping_matches_oracle_on_fresh_inputs(cuda_device):
    for mode, seed in (("people", 101), ("ties", 102), ("crowded", 103)):
        val, tag, ind = grouping_inputs(seed, 40, mode=mode)
        ans, num, scores = bottomup.group_by_tag(
            _t(val, cuda_device), _t(tag, cuda_device), _t(ind, cuda_device),
            synth.COCO_JOINT_ORDER, ignore_too_much=(mode == "crowded"))
        ans, num = ans.cpu().numpy(), num.cpu().numpy()
        for i in range(val.shape[0]):
            want = grouping.match_by_tag(val[i], tag[i], ind[i], synth.COCO_JOINT_ORDER,
                                         ignore_too_much=(mode == "crowded"))
            p = 0 if want.ndim == 1 else want.shape[0]
            assert num[i] == p
            if p:
                assert np.array_equal(ans[i, :p], want)


def test_grouping_empty_and_collisions(cuda_device):
    val = np.zeros((2, 17, 30), np.float32)
    tag = np.zeros((2, 17, 30, 1), np.float32)
    ind = np.zeros((2, 17, 30, 2), np.float32)
    val[1, 0, :2] = [0.9, 0.8]
    tag[1, 0, :2, 0] = [5.0, 5.0]          # equal keys: one group, later detection wins
    ind[1, 0, 0] = [1, 2]
    ind[1, 0, 1] = [3, 4]
    ans, num, _ = bottomup.group_by_tag(_t(val, cuda_device), _t(tag, cuda_device),
                                        _t(ind, cuda_device), synth.COCO_JOINT_ORDER)
    assert num.tolist() == [0, 1]
    assert ans[1, 0, 0].tolist() == [3.0, 4.0, np.float32(0.8), 5.0]


def test_grouping_more_than_128_people(cuda_device):
    """The reference's grouping is unbounded (match.py:63-113); here the capacity is a runtime
    parameter: 128 by default (num_groups = -1 flags the overflow), K * M can never overflow."""
    k, m = 17, 30
    rng = np.random.RandomState(5)
    val = rng.uniform(0.2, 1.0, (2, k, m)).astype(np.float32)
    ind = rng.randint(0, 128, (2, k, m, 2)).astype(np.float32)
    tag = np.zeros((2, k, m, 1), np.float32)
    # image 0: every detection far from every other one -> K * M people of one joint each;
    # image 1: even and odd joints in two disjoint tag ranges -> 2 * M people
    tag[0, :, :, 0] = (np.arange(k * m, dtype=np.float32) * 7.0).reshape(k, m)
    tag[1, :, :, 0] = np.arange(m, dtype=np.float32)[None] * 5.0 + \
        (np.arange(k)[:, None] % 2) * 1000.0 + rng.uniform(-0.3, 0.3, (k, m))
    args = [_t(x, cuda_device) for x in (val, tag, ind)]
    ans, num, _ = bottomup.group_by_tag(*args, synth.COCO_JOINT_ORDER)
    assert num.tolist() == [-1, 60]
    ans, num, scores = bottomup.group_by_tag(*args, synth.COCO_JOINT_ORDER, max_groups=k * m)
    assert ans.shape == (2, k * m, k, 4) and num.tolist() == [k * m, 60]
    ans, scores = ans.cpu().numpy(), scores.cpu().numpy()
    for i in range(2):
        want = grouping.match_by_tag(val[i], tag[i], ind[i], synth.COCO_JOINT_ORDER)
        assert np.array_equal(ans[i, :want.shape[0]], want), i
        assert scores[i, :want.shape[0]].tolist() == [float(x) for x in grouping.instance_scores(want)]
    # back-projection and refinement follow the capacity of `ans`
    center, scale = np.array([[64.0, 64.0]] * 2), np.array([[1.0, 1.0]] * 2)
    hw = np.array([[128.0, 128.0]] * 2)
    dev_ans, dev_num = _t(ans, cuda_device), _t(np.array([k * m, 60], np.int32), cuda_device)
    want = grouping.transform_keypoints([ans[0], ans[1, :60]], center, scale, hw)
    bottomup.transform_keypoints(dev_ans, dev_num, center, scale, hw)
    assert np.array_equal(dev_ans[0].cpu().numpy(), want[0])
    assert np.array_equal(dev_ans[1, :60].cpu().numpy(), want[1])
    with pytest.raises(ValueError, match="shared memory"):
        bottomup.group_by_tag(*args, synth.COCO_JOINT_ORDER, max_groups=2000)


def test_grouping_randomised_against_oracle(cuda_device):
    """60 random configurations (joints 3..32, max_num 1..64, up to 60 people, tag ties and
    equal keys, crowded and wide tag ranges, both norms, ignore_too_much, both capacities)
    against the oracle with scipy's own solver: the register solver (<= 32 and <= 64 groups)
    and the shared-memory one (more) all see traffic.  tests/stress/stress_grouping.py."""
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stress", "stress_grouping.py")
    spec = importlib.util.spec_from_file_location("stress_grouping", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run(60, seed=5, verbose=False) == 0


def test_transform_keypoints_matches_oracle(cuda_device):
    val, tag, ind = grouping_inputs(7, 6, mode="people")
    ans, num, _ = bottomup.group_by_tag(_t(val, cuda_device), _t(tag, cuda_device),
                                        _t(ind, cuda_device), synth.COCO_JOINT_ORDER)
    rng = np.random.RandomState(0)
    center = rng.uniform(100, 400, (6, 2))
    scale = rng.uniform(1, 3.2, (6, 2))
    hw = np.tile(np.array([[256.0, 256.0]]), (6, 1))
    before = [ans[i, :int(num[i])].cpu().numpy() for i in range(6)]
    want = grouping.transform_keypoints(before, center, scale, hw, pixel_std=200.0)
    bottomup.transform_keypoints(ans, num, center, scale, hw, pixel_std=200.0)
    for i in range(6):
        assert np.array_equal(ans[i, :int(num[i])].cpu().numpy(), want[i])


@pytest.mark.parametrize("refine", [False, True])
def test_bottomup_inferencer_end_to_end(cuda_device, refine):
    """decode -> group -> score -> [refine missing joints] -> back-project through the
    registry classes."""
    from oracle import refine_missing as rm

    d = synth.bottomup_outputs(2, 17, 32, 32, mask_hw=(128, 128), seed=11, max_people=4)
    dev = cuda_device
    out = [_t(d["out0"], dev), _t(d["out1"], dev)]
    cfg = dict(has_heatmap_output=True, hflip_tta=False, joint_order=synth.COCO_JOINT_ORDER,
               vis_thr=0.1, ignore_too_much=False, use_rounded_norm=True, tag_thr=1.0,
               pixel_std=200.0, downsample_scale=2, refine_missing_joint=refine,
               flip_pairs=synth.COCO_FLIP_PAIRS)
    dec = mp.create_decoder("bottomup_heatmap_ae", use_nms=True, nms_kernel=3)
    inf = mp.create_inferencer(lambda image: out, "bottomup_heatmap_ae", config=cfg, decoder=dec)
    center = np.array([[64.0, 64.0], [60.0, 70.0]])
    scale = np.array([[0.64, 0.64], [0.7, 0.7]])
    shape = np.array([[128.0, 128.0], [128.0, 128.0]])
    data = dict(image=None, mask=_t(d["mask"], dev), center=center, scale=scale,
                image_shape=shape, image_file=["a.jpg", "b.jpg"])
    records = inf([data])
    val_k, tag_k, ind_k, raw, tagging = bd.decode([d["out0"], d["out1"]], d["mask"], use_nms=True,
                                                  nms_kernel=3, max_num=30)
    people = [grouping.match_by_tag(val_k[i], tag_k[i], ind_k[i], synth.COCO_JOINT_ORDER)
              for i in range(2)]
    people = [p if p.ndim == 3 else np.zeros((0, 17, 4), np.float32) for p in people]
    scores = [[y[:, 2].mean() for y in p] for p in people]  # before the refinement (:153-156)
    if refine:
        people = [np.stack([rm.refine_missing(raw[i], tagging[i], kp) for kp in p])
                  if len(p) else p for i, p in enumerate(people)]
    want = grouping.transform_keypoints(people, center, scale, shape / 2, pixel_std=200.0)
    assert len(records) == 2
    for rec, w, sc in zip(records, want, scores):
        assert np.array_equal(rec["pred"], w)
        assert rec["score"] == sc


def test_bottomup_inferencer_more_than_128_people(cuda_device):
    """Tags spread over a wide range: nearly every detection is its own person (> 128); the
    inferencer groups again with room for K * M people instead of failing."""
    from oracle import refine_missing as rm

    rng = np.random.RandomState(3)
    n, k, h = 1, 17, 32
    out0 = np.concatenate([rng.uniform(0, 1, (n, k, h, h)),
                           rng.uniform(-500, 500, (n, k, h, h))], 1).astype(np.float32)
    out1 = rng.uniform(0, 1, (n, k, 2 * h, 2 * h)).astype(np.float32)
    mask = np.ones((n, 4 * h, 4 * h), np.uint8)
    dev = cuda_device
    cfg = dict(has_heatmap_output=True, hflip_tta=False, joint_order=synth.COCO_JOINT_ORDER,
               vis_thr=0.1, ignore_too_much=False, use_rounded_norm=True, tag_thr=1.0,
               pixel_std=200.0, downsample_scale=2, refine_missing_joint=True,
               flip_pairs=synth.COCO_FLIP_PAIRS)
    dec = mp.create_decoder("bottomup_heatmap_ae", use_nms=True, nms_kernel=3)
    inf = mp.create_inferencer(lambda image: [_t(out0, dev), _t(out1, dev)],
                               "bottomup_heatmap_ae", config=cfg, decoder=dec)
    center, scale = np.array([[64.0, 64.0]]), np.array([[0.64, 0.64]])
    shape = np.array([[128.0, 128.0]])
    records = inf([dict(image=None, mask=_t(mask, dev), center=center, scale=scale,
                        image_shape=shape, image_file=["a.jpg"])])
    val_k, tag_k, ind_k, raw, tagging = bd.decode([out0, out1], mask, use_nms=True, nms_kernel=3,
                                                  max_num=30)
    people = grouping.match_by_tag(val_k[0], tag_k[0], ind_k[0], synth.COCO_JOINT_ORDER)
    assert people.shape[0] > 128
    scores = [y[:, 2].mean() for y in people]
    people = np.stack([rm.refine_missing(raw[0], tagging[0], kp) for kp in people])
    want = grouping.transform_keypoints([people], center, scale, shape / 2, pixel_std=200.0)
    assert np.array_equal(records[0]["pred"], want[0])
    assert records[0]["score"] == scores


# ------------------------------------------------------------------ N1: bottom-up encode
def test_bottomup_encode_matches_reference_golden(cuda_device, golden):
    from oracle import gen_golden_bottomup as ggb

    g = golden("bottomup_encode_ref.npz")
    for case, (sizes, m, tpj, seed) in enumerate(ggb.BOTTOMUP_ENCODE_CASES):
        kps = ggb.bottomup_people(seed, m, 17, sizes)
        cfg = dict(image_size=[512, 512], max_image_size=[832, 512], heatmap_sizes=sizes,
                   flip_pairs=[[1, 2]], pixel_std=200.0, tag_per_joint=tpj)
        t = mp.create_transform("bottomup_generate_target", is_train=True, config=cfg, sigma=2.0,
                                max_num=30)
        batch = _t(np.stack(kps)[None], cuda_device)
        target, tag_ind = t.encode_batch(batch)
        want_t, want_g = g[f"target_{case}"], g[f"tag_ind_{case}"]
        assert target.shape[1:] == want_t.shape and tag_ind.shape[1:] == want_g.shape
        assert np.abs(target[0].cpu().numpy() - want_t).max() <= 1e-5, case
        assert np.array_equal(tag_ind[0].cpu().numpy(), want_g), case
        # per-sample calling convention (the reference's transform(state))
        out = t.transform(dict(keypoints=kps))
        assert np.array_equal(out["tag_ind"], want_g)
        assert np.abs(out["target"] - want_t).max() <= 1e-5


def test_bottomup_encode_full_size_batch(cuda_device):
    """BASELINE config 4 shapes (128^2 + 256^2, 17 joints), batch of 16 images with different
    people counts padded to M = 30; every image against the oracle."""
    from oracle import bottomup_encode as be
    from oracle import gen_golden_bottomup as ggb

    sizes = [[128, 128], [256, 256]]
    n, m = 16, 30
    batch = np.zeros((n, 2, m, 17, 3), np.float32)
    people = []
    for i in range(n):
        cnt = [0, 1, 5, 12, 30][i % 5]
        kps = ggb.bottomup_people(50 + i, cnt, 17, sizes)
        people.append(kps)
        for s in range(2):
            batch[i, s, :cnt] = kps[s]
    target, tag_ind = bottomup.encode_targets(_t(batch, cuda_device), sizes)
    target, tag_ind = target.cpu().numpy(), tag_ind.cpu().numpy()
    for i in range(n):
        want_t, want_g = be.encode(people[i], sizes)
        assert np.abs(target[i] - want_t).max() <= 1e-5, i
        assert np.array_equal(tag_ind[i], want_g), i
    with pytest.raises(ValueError, match="exeeds the maximum num"):
        bottomup.encode_targets(torch.zeros(1, 2, 31, 17, 3, device=cuda_device), sizes)


# ------------------------------------------------------------------ N3: refine_missing
def _refine_batch(dev, cases):
    """cases: list of (heat [K,H,W], tag [K,H,W,1], kps [P,K,4]) of one map size."""
    n = len(cases)
    k, h, w = cases[0][0].shape
    g = bottomup._lib.PC_MAX_GROUPS
    heat = np.stack([c[0] for c in cases])
    tagm = np.stack([c[1] for c in cases])
    ans = np.zeros((n, g, k, 4), np.float32)
    num = np.zeros(n, np.int32)
    for i, c in enumerate(cases):
        num[i] = c[2].shape[0]
        ans[i, :num[i]] = c[2]
    out = bottomup.refine_missing(_t(heat, dev), _t(tagm, dev), _t(ans, dev), _t(num, dev))
    return out.cpu().numpy(), num


def test_refine_missing_matches_reference_golden(cuda_device, golden):
    g = golden("refine_missing_ref.npz")
    cases = [(g[f"heat_{s}"], g[f"tag_{s}"], g[f"kps_{s}"]) for s in range(4)]
    got, num = _refine_batch(cuda_device, cases)
    for s in range(4):
        assert np.array_equal(got[s, :num[s]], g[f"refined_{s}"]), s


def test_refine_missing_full_size_and_inferencer(cuda_device):
    """256 x 256 maps, 20 people (three chunks of 8), and the inferencer flag end to end."""
    from oracle import gen_golden_bottomup as ggb
    from oracle import refine_missing as rm

    cases = [ggb.refine_inputs(40 + i, h=256, w=256, people=p) for i, p in enumerate((20, 0, 9))]
    got, num = _refine_batch(cuda_device, cases)
    for i, (heat, tagm, kps) in enumerate(cases):
        for p in range(num[i]):
            assert np.array_equal(got[i, p], rm.refine_missing(heat, tagm, kps[p])), (i, p)
