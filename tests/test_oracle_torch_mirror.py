"""Second opinion on the two PARITY-UNPINNED oracles (oracle/topdown_decode.py and
oracle/bottomup_decode.py).

MindSpore cannot be installed here, so the decoder graphs
(mindpose/models/decoders/top_down_decoder.py:72-215,
mindpose/models/decoders/bottom_up_decoder.py:67-203,
mindpose/engine/inferencer/topdown_inferencer.py:165-187) cannot be executed.  What CAN
be executed is the same graph spelled with another tensor library's operators whose
documented semantics match the MindSpore 1.x operators the reference calls:

    ops.max(axis, keep_dims)          -> torch.max(dim, keepdim)
    tensor_scatter_elements / masked_select -> Tensor.scatter_ / torch.masked_select
    ops.conv2d(group=K, pad_mode="same")    -> F.conv2d(groups=K, padding=5)
    ops.clip_by_value / ops.log / ops.pad   -> torch.clamp / torch.log / F.pad
    ops.MatrixInverse / ops.Einsum          -> torch.linalg.inv / torch.einsum
    nn.MaxPool2d(k, pad_mode="same")        -> F.max_pool2d(k, 1, k // 2)  (odd k)
    ops.top_k                               -> torch.topk
    ops.ResizeNearestNeighbor               -> F.interpolate(mode="nearest")
    ops.ResizeBilinear (legacy asymmetric)  -> F.grid_sample on the legacy source grid
    ops.gather_elements                     -> torch.gather

The torch graph below is written operator by operator after the reference source (line
numbers in the comments) and shares no code with the numpy restatement, so agreement
checks the restatement's reading of the graph (op order, slicing, index arithmetic,
the masked_select pairing quirk), not just its arithmetic.  It does not pin MindSpore's
backend-defined choices (tie order of max / top_k, conv accumulation order): those stay
"parity unpinned" and the tests avoid ties / use the stated tolerance where they matter.
This file never touches the CUDA path.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from mindpose_b200 import synth
from oracle import bottomup_decode as obd
from oracle import topdown_decode as otd


# ---------------------------------------------------------------------------
# top-down (top_down_decoder.py, topdown_inferencer.py)
# ---------------------------------------------------------------------------
def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def td_flip_average(heatmap, flipped, flip_index, shift_heatmap):
    back = flipped[:, torch.as_tensor(flip_index, dtype=torch.long)]   # inferencer :180
    back = torch.flip(back, dims=[3])                                  # :181
    if shift_heatmap:                                                  # :185 (RHS read first)
        back = back.clone()
        back[..., 1:] = back[..., :-1].clone()
    return (heatmap + back) * 0.5                                      # :176


def td_get_max_preds(heatmap):
    n, k, _, w = heatmap.shape
    flat = heatmap.reshape(n, k, -1)                                   # decoder :101
    maxvals, idx = torch.max(flat, dim=2, keepdim=True)                # :102
    mask = torch.zeros(flat.shape, dtype=torch.bool)
    mask.scatter_(2, idx, torch.ones(idx.shape, dtype=torch.bool))     # :105-108
    mask = mask.reshape(n, k, -1, w)
    preds = idx.repeat(1, 1, 2).to(torch.float32)                      # :111
    preds[:, :, 0] = preds[:, :, 0] % w                                # :113
    preds[:, :, 1] = torch.floor(preds[:, :, 1] / w)                   # :114
    return preds, maxvals, mask


def td_shift_coordinate(coords, heatmap, mask):
    n, k = coords.shape[:2]
    dx = torch.zeros_like(heatmap)
    dy = torch.zeros_like(heatmap)
    dx[:, :, :, 1:-1] = heatmap[:, :, :, 2:] - heatmap[:, :, :, :-2]   # :128
    dy[:, :, 1:-1, :] = heatmap[:, :, 2:, :] - heatmap[:, :, :-2, :]   # :129
    ox = torch.masked_select(torch.sign(dx), mask).reshape(n, k) * 0.25  # :133-136
    oy = torch.masked_select(torch.sign(dy), mask).reshape(n, k) * 0.25
    coords = coords.clone()
    coords[..., 0] += ox
    coords[..., 1] += oy
    return coords


def td_gaussian_kernel(kernel_size):
    sigma = 0.3 * ((kernel_size - 1) * 0.5 - 1) + 0.8                  # :208
    xs = np.arange(-(kernel_size - 1) // 2, (kernel_size - 1) // 2 + 1, 1)
    ys = xs[:, None]
    kern = np.exp(-(xs ** 2 + ys ** 2) / (2 * sigma ** 2))
    kern = kern / kern.sum()
    return torch.tensor(kern[None, None], dtype=torch.float32)         # :213-214


def td_dark_refine(coords, heatmap, kernel_size):
    n, k, h, w = heatmap.shape
    kern = td_gaussian_kernel(kernel_size).repeat(k, 1, 1, 1)          # :174
    hm = F.conv2d(heatmap, kern, groups=k, padding=(kernel_size - 1) // 2)  # :175
    hm = torch.clamp(hm, 0.001, 50)                                    # :176
    hm = torch.log(hm)                                                 # :177
    hm = F.pad(hm, (1, 1, 1, 1))                                       # :178
    hm = hm.flatten()
    index = coords[..., 0] + 1 + (coords[..., 1] + 1) * (w + 2)        # :181 (float32)
    index = index + (w + 2) * (h + 2) * torch.arange(0, n * k, 1).reshape(-1, k)  # :182
    index = index.to(torch.int32).reshape(-1, 1).long()                # :183
    i_ = hm[index]
    ix1 = hm[index + 1]
    iy1 = hm[index + w + 2]
    ix1y1 = hm[index + w + 3]
    ix1_y1_ = hm[index - w - 3]
    ix1_ = hm[index - 1]
    iy1_ = hm[index - 2 - w]
    dx = 0.5 * (ix1 - ix1_)                                            # :192
    dy = 0.5 * (iy1 - iy1_)
    derivative = torch.cat([dx, dy], dim=1).reshape(n, k, 2, 1)
    dxx = ix1 - 2 * i_ + ix1_                                          # :197
    dyy = iy1 - 2 * i_ + iy1_
    dxy = 0.5 * (ix1y1 - ix1 - iy1 + i_ + i_ - ix1_ - iy1_ + ix1_y1_)  # :199
    hessian = torch.cat([dxx, dxy, dxy, dyy], dim=1).reshape(n, k, 2, 2)
    hessian = torch.linalg.inv(hessian + torch.eye(2) * 1e-7)          # :203
    return coords - torch.einsum("ijmn,ijnk->ijmk", hessian, derivative).squeeze(-1)  # :204


def td_transform_preds(coords, center, scale, hw, pixel_std, use_udp):
    scale = scale * pixel_std                                          # :152
    if use_udp:
        sx = scale[:, 0:1] / (hw[1] - 1.0)
        sy = scale[:, 1:2] / (hw[0] - 1.0)
    else:
        sx = scale[:, 0:1] / hw[1]
        sy = scale[:, 1:2] / hw[0]
    out = torch.ones_like(coords)
    out[:, :, 0] = coords[:, :, 0] * sx + center[:, 0:1] - scale[:, 0:1] * 0.5   # :162-164
    out[:, :, 1] = coords[:, :, 1] * sy + center[:, 1:2] - scale[:, 1:2] * 0.5
    return out


def td_decode(heatmap, center, scale, score, pixel_std=200.0, to_original=True,
              shift_coordinate=False, use_udp=False, dark_udp_refine=False, kernel_size=11):
    n = heatmap.shape[0]
    coords, maxvals, mask = td_get_max_preds(heatmap)                  # :77
    if shift_coordinate:
        coords = td_shift_coordinate(coords, heatmap, mask)
    elif dark_udp_refine:
        coords = td_dark_refine(coords, heatmap, kernel_size)
    if to_original:
        coords = td_transform_preds(coords, center, scale, heatmap.shape[2:], pixel_std,
                                    use_udp)
    preds = torch.zeros((n, coords.shape[1], 3))
    boxes = torch.zeros((n, 6))
    preds[:, :, 0:2] = coords[:, :, 0:2]
    preds[:, :, 2:3] = maxvals
    boxes[:, 0:2] = center[:, 0:2]
    boxes[:, 2:4] = scale[:, 0:2]
    boxes[:, 4] = torch.prod(scale * pixel_std, dim=1)                 # :90
    boxes[:, 5] = score
    return preds.numpy(), boxes.numpy()


@pytest.mark.parametrize("h,w", [(64, 48), (96, 72), (17, 23)])
@pytest.mark.parametrize("use_udp", [False, True])
@pytest.mark.parametrize("shift", [False, True])
def test_topdown_plain_and_quarter_shift_agree_bit_for_bit(h, w, use_udp, shift):
    n, k = 6, 17
    maps = synth.noise_heatmaps(n, k, h, w, seed=5)   # what the reference's tests feed
    c, s, sc = synth.crop_geometry(n, seed=5)
    want_p, want_b = otd.decode(maps, c, s, sc, shift_coordinate_flag=shift, use_udp=use_udp)
    got_p, got_b = td_decode(_t(maps), _t(c), _t(s), _t(sc), shift_coordinate=shift,
                             use_udp=use_udp)
    assert np.array_equal(got_p, want_p)
    assert np.array_equal(got_b, want_b)


@pytest.mark.parametrize("shift_heatmap", [False, True])
def test_topdown_flip_average_agrees_bit_for_bit(shift_heatmap):
    n, k, h, w = 5, 17, 64, 48
    maps, _ = synth.blob_heatmaps(n, k, h, w, seed=2)
    flipped = synth.flipped_pair(maps, seed=2)
    fidx = synth.flip_index()
    want = otd.flip_average(maps, flipped, fidx, shift_heatmap)
    got = td_flip_average(_t(maps), _t(flipped), fidx, shift_heatmap).numpy()
    assert np.array_equal(got, want)
    c, s, sc = synth.crop_geometry(n, seed=2)
    want_p, want_b = otd.decode_with_flip(maps, flipped, fidx, c, s, sc, shift_heatmap,
                                          shift_coordinate_flag=True)
    got_p, got_b = td_decode(td_flip_average(_t(maps), _t(flipped), fidx, shift_heatmap),
                             _t(c), _t(s), _t(sc), shift_coordinate=True)
    assert np.array_equal(got_p, want_p) and np.array_equal(got_b, want_b)


@pytest.mark.parametrize("h,w,use_udp", [(64, 48, False), (96, 72, True)])
def test_topdown_dark_agrees_within_the_stated_tolerance(h, w, use_udp):
    """The blur's float32 accumulation order is backend-defined (the oracle states its
    order of record); everything else is op for op.  Heat-map-pixel coordinates must
    agree to 1e-4 px (north_star's bar for refined coordinates)."""
    n, k = 8, 17
    maps, _ = synth.blob_heatmaps(n, k, h, w, seed=9)
    c, s, sc = synth.crop_geometry(n, seed=9)
    want_p, _ = otd.decode(maps, c, s, sc, dark_udp_refine_flag=True, use_udp=use_udp,
                           to_original=False)
    got_p, _ = td_decode(_t(maps), _t(c), _t(s), _t(sc), dark_udp_refine=True, use_udp=use_udp,
                         to_original=False)
    assert np.array_equal(got_p[..., 2], want_p[..., 2])
    assert np.abs(got_p[..., :2] - want_p[..., :2]).max() <= 1e-4
    # and the refinement moved the integer peak (the test is not vacuous)
    plain, _ = otd.decode(maps, c, s, sc, to_original=False)
    assert np.abs(want_p[..., :2] - plain[..., :2]).max() > 0.2


def test_topdown_dark_peak_on_the_border_reads_the_zero_pad():
    """A peak in row 0 / column 0 takes its outside neighbours from the zero pad of the
    LOG map (:178), not from log(0.001)."""
    n, k, h, w = 2, 3, 16, 12
    rng = np.random.RandomState(0)
    maps = rng.uniform(0, 0.02, (n, k, h, w)).astype(np.float32)
    maps[:, 0, 0, 0] = 0.9
    maps[:, 1, h - 1, w - 1] = 0.8
    maps[:, 2, 0, 5] = 0.7
    c, s, sc = synth.crop_geometry(n, seed=1)
    want_p, _ = otd.decode(maps, c, s, sc, dark_udp_refine_flag=True, to_original=False)
    got_p, _ = td_decode(_t(maps), _t(c), _t(s), _t(sc), dark_udp_refine=True,
                         to_original=False)
    assert np.abs(got_p - want_p).max() <= 1e-4


# ---------------------------------------------------------------------------
# bottom-up (bottom_up_decoder.py)
# ---------------------------------------------------------------------------
def bu_resize_bilinear_legacy(x, out_h, out_w):
    """ops.ResizeBilinear(size), align_corners=False, half_pixel_centers=False: source
    coordinate = dst * in / out, clamped at the last row / column -- expressed as a
    grid_sample (align_corners=True, border padding) over that source grid."""
    _, _, in_h, in_w = x.shape
    ys = torch.arange(out_h, dtype=torch.float64) * (in_h / out_h)
    xs = torch.arange(out_w, dtype=torch.float64) * (in_w / out_w)
    gy = (2 * ys / (in_h - 1) - 1).clamp(max=1.0)
    gx = (2 * xs / (in_w - 1) - 1).clamp(max=1.0)
    grid = torch.stack(torch.meshgrid(gx, gy, indexing="xy"), dim=-1)
    grid = grid[None].expand(x.shape[0], -1, -1, -1)
    out = F.grid_sample(x.double(), grid, mode="bilinear", padding_mode="border",
                        align_corners=True)
    return out.float()


def bu_decode(model_output, mask, num_joints=17, num_stages=2, with_ae_loss=(True, False),
              use_nms=False, nms_kernel=5, max_num=30, shift_coordinate=False,
              aggregated=None):
    heat, tag = [], []
    for i in range(num_stages):                                        # :97-101
        heat.append(model_output[i][:, :num_joints])
        if with_ae_loss[i]:
            tag.append(model_output[i][:, num_joints:])
    mask = mask[:, None, ...]                                          # :110
    if num_stages > 1:                                                 # :129-138
        base = heat[-1].clone()
        hh, ww = base.shape[2:]
        for i in range(num_stages - 1):
            base += bu_resize_bilinear_legacy(heat[i], hh, ww)
        base /= num_stages
    else:
        base = heat[0]
    hh, ww = base.shape[2:]
    tags = torch.stack([bu_resize_bilinear_legacy(t, hh, ww) for t in tag], dim=-1)  # :120-122
    m = F.interpolate(mask.to(base.dtype), size=(hh, ww), mode="nearest").bool()  # :125-126
    base = base.masked_fill(~m, 0)                                     # :127
    if aggregated is not None:   # continue from the oracle's bits (rounding of the lerp differs)
        base, tags = aggregated
    raw = base.clone()                                                 # :77
    if use_nms:                                                        # :173-178
        pooled = F.max_pool2d(base, nms_kernel, 1, nms_kernel // 2)
        base = base * torch.eq(pooled, base).to(base.dtype)
    n, k = base.shape[:2]
    flat = base.reshape(n, k, -1)
    val_k, ind = torch.topk(flat, max_num, dim=2)                      # :147
    sel = torch.zeros(flat.shape, dtype=torch.bool)
    sel.scatter_(2, ind, torch.ones(ind.shape, dtype=torch.bool))      # :150-153
    sel = sel.reshape(n, k, hh, ww)
    tflat = tags.reshape(n, tags.shape[1], ww * hh, -1)                # :156
    tag_k = torch.stack([torch.gather(tflat[..., i], 2, ind) for i in range(tflat.shape[3])],
                        dim=3)                                         # :160-163
    ind_k = torch.stack((ind % ww, ind // ww), dim=3).to(val_k.dtype)  # :165-169
    if shift_coordinate:                                               # :180-203
        dx = torch.zeros_like(raw)
        dy = torch.zeros_like(raw)
        dx[:, :, :, 1:-1] = raw[:, :, :, 2:] - raw[:, :, :, :-2]
        dy[:, :, 1:-1, :] = raw[:, :, 2:, :] - raw[:, :, :-2, :]
        ox = torch.masked_select(torch.sign(dx), sel).reshape(n, k, -1) * 0.25
        oy = torch.masked_select(torch.sign(dy), sel).reshape(n, k, -1) * 0.25
        ind_k = ind_k.clone()
        ind_k[..., 0] += ox
        ind_k[..., 1] += oy
    return val_k.numpy(), tag_k.numpy(), ind_k.numpy(), raw.numpy(), tags.numpy()


def _bu_case(seed, n=2, h0=24, w0=20, mask_hw=(96, 80)):
    d = synth.bottomup_outputs(n, 17, h0, w0, mask_hw=mask_hw, seed=seed, max_people=3)
    return [d["out0"], d["out1"]], d["mask"]


def test_bottomup_aggregation_geometry_matches_the_legacy_resize():
    """Same source pixels and weights as the restatement (the lerp is evaluated in another
    order, so the comparison is to 1e-6, not bit for bit); the mask goes through
    F.interpolate(mode="nearest"), which floors dst * in / out like
    ResizeNearestNeighbor."""
    outs, mask = _bu_case(0)
    want = obd.decode(outs, mask, use_nms=False, max_num=30)
    got = bu_decode([_t(o) for o in outs], _t(mask), use_nms=False, max_num=30)
    assert np.abs(got[3] - want[3]).max() <= 1e-6      # heatmap_raw
    assert np.abs(got[4] - want[4]).max() <= 2e-6      # tagging_heatmap
    assert np.array_equal(got[3] == 0, want[3] == 0)   # the mask zeroes the same pixels
    # an upsampling factor other than 2 (the legacy mapping is not the half-pixel one)
    x = np.random.RandomState(1).rand(1, 2, 5, 7).astype(np.float32)
    assert np.abs(bu_resize_bilinear_legacy(_t(x), 15, 21).numpy()
                  - obd.resize_bilinear_legacy(x, 15, 21)).max() <= 1e-6
    half_pixel = F.interpolate(_t(x), size=(15, 21), mode="bilinear", align_corners=False)
    assert np.abs(half_pixel.numpy() - obd.resize_bilinear_legacy(x, 15, 21)).max() > 1e-2


@pytest.mark.parametrize("use_nms,nms_kernel", [(False, 5), (True, 3), (True, 5)])
@pytest.mark.parametrize("shift", [False, True])
def test_bottomup_nms_topk_gather_and_shift_agree_bit_for_bit(use_nms, nms_kernel, shift):
    """From the same aggregated maps on: pool, equality mask, top-k, tag gather, (x, y),
    and the masked_select pairing of the quarter-pixel shift.  Compared where the
    selection is not a tie (torch.topk's tie order is its own)."""
    outs, mask = _bu_case(3)
    want = obd.decode(outs, mask, use_nms=use_nms, nms_kernel=nms_kernel, max_num=30,
                      shift_coordinate=shift)
    agg = (_t(want[3]), _t(want[4]))
    got = bu_decode([_t(o) for o in outs], _t(mask), use_nms=use_nms, nms_kernel=nms_kernel,
                    max_num=30, shift_coordinate=shift, aggregated=agg)
    assert np.array_equal(got[0], want[0])             # values: ties cannot change them
    v = want[0]
    n, k, m = v.shape
    # a rank is unambiguous if its value differs from both neighbours and from the value
    # that just missed the cut (rank m + 1)
    nxt = obd.decode(outs, mask, use_nms=use_nms, nms_kernel=nms_kernel, max_num=m + 1)[0]
    uniq = np.ones_like(v, dtype=bool)
    uniq[..., 1:] &= v[..., 1:] != v[..., :-1]
    uniq[..., :-1] &= v[..., :-1] != v[..., 1:]
    uniq[..., -1] &= v[..., -1] != nxt[..., m]
    assert uniq.mean() > 0.3
    if not shift:
        assert np.array_equal(got[2][uniq], want[2][uniq])
        assert np.array_equal(got[1][uniq], want[1][uniq])
    else:
        # the shift pairs offsets in SPATIAL order with entries in RANK order: it only
        # depends on the selected SET, which is tie-free when the cut is
        cut_ok = (v[..., -1] != nxt[..., m])
        assert cut_ok.mean() > 0.3
        assert np.array_equal(got[2][cut_ok][uniq[cut_ok]], want[2][cut_ok][uniq[cut_ok]])
        # and the quirk is real: pairing in rank order would give another answer
        plain = obd.decode(outs, mask, use_nms=use_nms, nms_kernel=nms_kernel, max_num=30)[2]
        assert not np.array_equal(want[2] - plain, np.zeros_like(plain))
