"""Parity of the CUDA top-down codec against the oracle (B200 only).

Bars (BASELINE.json north_star): argmax indices / maxvals bit-exact; refined
coordinates within 1e-4 px; encoded targets within 1e-5; warped pixels bit-exact.
"""
import numpy as np
import pytest
import torch

import mindpose_b200 as mp
from mindpose_b200 import codec, synth
from oracle import affine, topdown_decode, topdown_encode, warp

pytestmark = pytest.mark.gpu

COORD_TOL = 1e-4
TARGET_TOL = 1e-5


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


# --------------------------------------------------------------------- decode
def _decode_case(dev, maps, flipped, n, h, w, seed, *, flip, shift_heatmap, shift_coordinate,
                 dark, use_udp, to_original=True, kernel_size=11):
    center, scale, score = synth.crop_geometry(n, seed=seed)
    fidx = synth.flip_index()
    kw = dict(to_original=to_original, shift_coordinate_flag=shift_coordinate, use_udp=use_udp,
              dark_udp_refine_flag=dark, kernel_size=kernel_size)
    if flip:
        want_p, want_b = topdown_decode.decode_with_flip(maps, flipped, fidx, center, scale, score,
                                                         shift_heatmap=shift_heatmap, **kw)
    else:
        want_p, want_b = topdown_decode.decode(maps, center, scale, score, **kw)
    dec = mp.create_decoder("topdown_heatmap", to_original=to_original,
                            shift_coordinate=shift_coordinate, use_udp=use_udp,
                            dark_udp_refine=dark, kernel_size=kernel_size)
    if flip:
        got_p, got_b = dec.decode_flip_pair(_t(maps, dev), _t(flipped, dev), fidx, _t(center, dev),
                                            _t(scale, dev), _t(score, dev),
                                            shift_heatmap=shift_heatmap)
    else:
        got_p, got_b = dec(_t(maps, dev), _t(center, dev), _t(scale, dev), _t(score, dev))
    torch.cuda.synchronize()
    return got_p.cpu().numpy(), got_b.cpu().numpy(), want_p, want_b


@pytest.mark.parametrize("h,w", [(64, 48), (96, 72), (48, 64)])
@pytest.mark.parametrize("flip,shift_heatmap", [(False, False), (True, False), (True, True)])
@pytest.mark.parametrize("mode", ["plain", "shift", "dark", "dark_udp"])
def test_decode_blobs(cuda_device, h, w, flip, shift_heatmap, mode):
    n = 33
    maps, _ = synth.blob_heatmaps(n, 17, h, w, seed=h + w)
    flipped = synth.flipped_pair(maps, seed=h)
    got_p, got_b, want_p, want_b = _decode_case(
        cuda_device, maps, flipped, n, h, w, seed=3, flip=flip, shift_heatmap=shift_heatmap,
        shift_coordinate=(mode == "shift"), dark=mode.startswith("dark"),
        use_udp=(mode == "dark_udp"))
    assert np.array_equal(got_p[..., 2], want_p[..., 2])  # maxvals bit-exact
    assert np.array_equal(got_b, want_b)
    if mode in ("plain", "shift"):
        assert np.array_equal(got_p, want_p)  # exact-order fp32: bit-exact
    else:
        assert np.abs(got_p[..., :2] - want_p[..., :2]).max() <= COORD_TOL


@pytest.mark.parametrize("h,w", [(17, 13), (31, 27), (64, 49), (5, 3), (30, 22)])
@pytest.mark.parametrize("flip,shift_heatmap", [(False, False), (True, True)])
@pytest.mark.parametrize("mode", ["plain", "shift", "dark"])
def test_decode_any_map_size(cuda_device, h, w, flip, shift_heatmap, mode):
    """Map sizes the reference accepts (any H, W: top_down_decoder.py:96-116) whose planes are
    not 16-byte aligned in HBM (H * W % 4 != 0, or W % 4 != 0): the producer warp copies the
    planes itself instead of bulk-copying them; same results."""
    n = 9
    maps, _ = synth.blob_heatmaps(n, 17, h, w, seed=h + w)
    flipped = synth.flipped_pair(maps, seed=h)
    got_p, got_b, want_p, want_b = _decode_case(
        cuda_device, maps, flipped, n, h, w, seed=3, flip=flip, shift_heatmap=shift_heatmap,
        shift_coordinate=(mode == "shift"), dark=(mode == "dark"), use_udp=False)
    assert np.array_equal(got_p[..., 2], want_p[..., 2])
    assert np.array_equal(got_b, want_b)
    if mode != "dark":
        assert np.array_equal(got_p, want_p)
    else:
        assert np.abs(got_p[..., :2] - want_p[..., :2]).max() <= COORD_TOL


def test_decode_unaligned_base_pointer(cuda_device):
    """A heat-map view that starts 4 bytes into a buffer (planes not 16-byte aligned although
    H * W % 4 == 0) takes the manual copy too."""
    dev = cuda_device
    n, k, h, w = 5, 17, 64, 48
    maps, _ = synth.blob_heatmaps(n, k, h, w, seed=4)
    center, scale, score = synth.crop_geometry(n, seed=4)
    want_p, want_b = topdown_decode.decode(maps, center, scale, score, shift_coordinate_flag=True)
    buf = torch.zeros(maps.size + 1, device=dev)
    view = buf[1:].view(n, k, h, w)
    view.copy_(_t(maps, dev))
    assert view.data_ptr() % 16 == 4
    p = codec.make_decode_params(k, h, w, shift_coordinate=True)
    got_p, got_b = codec.topdown_decode(view, _t(center, dev), _t(scale, dev), _t(score, dev),
                                        params=p)
    assert np.array_equal(got_p.cpu().numpy(), want_p) and np.array_equal(got_b.cpu().numpy(), want_b)


@pytest.mark.parametrize("flip", [False, True])
def test_decode_noise_maps_indices_bit_exact(cuda_device, flip):
    """What the reference's own tests feed (uniform noise): argmax, maxval and the
    quarter-offset are exact; DARK offsets are ill-conditioned there and are not compared."""
    n, h, w = 64, 64, 48
    maps = synth.noise_heatmaps(n, 17, h, w, seed=0)
    flipped = synth.noise_heatmaps(n, 17, h, w, seed=1)
    for to_original in (False, True):
        got_p, got_b, want_p, want_b = _decode_case(
            cuda_device, maps, flipped, n, h, w, seed=0, flip=flip, shift_heatmap=True,
            shift_coordinate=True, dark=False, use_udp=False, to_original=to_original)
        assert np.array_equal(got_p, want_p)
        assert np.array_equal(got_b, want_b)


def test_decode_ties_take_lowest_index_and_border_peaks(cuda_device):
    n, k, h, w = 3, 17, 64, 48
    maps = np.zeros((n, k, h, w), np.float32)
    maps[0, :, 10, 7] = 1.0
    maps[0, :, 50, 3] = 1.0           # duplicate maximum later in the plane
    maps[1, :, 0, 0] = 0.7            # corner peak: DARK reads the zero padding of the log map
    maps[1, 5, 63, 47] = 0.9
    maps[2] = 0.25                    # constant plane: argmax 0
    center, scale, score = synth.crop_geometry(n, seed=5)
    for dark in (False, True):
        want_p, want_b = topdown_decode.decode(maps, center, scale, score, to_original=False,
                                               dark_udp_refine_flag=dark)
        dec = mp.create_decoder("topdown_heatmap", to_original=False, dark_udp_refine=dark)
        got_p, got_b = dec(_t(maps, cuda_device), _t(center, cuda_device), _t(scale, cuda_device),
                           _t(score, cuda_device))
        got_p = got_p.cpu().numpy()
        assert np.array_equal(got_p[..., 2], want_p[..., 2])
        assert np.allclose(got_p, want_p, rtol=0, atol=COORD_TOL)
    assert got_p.shape == (n, k, 3)


def test_decode_reference_test_shapes(cuda_device):
    """tests/models/decoders/test_top_down_decoder.py: [8,17,48,64] -> (8,17,3), (8,6)."""
    dev = cuda_device
    for kwargs in (dict(), dict(shift_coordinate=True), dict(use_udp=True, dark_udp_refine=True)):
        dec = mp.create_decoder("topdown_heatmap", **kwargs)
        hm = torch.rand(8, 17, 48, 64, device=dev)
        p, b = dec(hm, torch.rand(8, 2, device=dev) * 400, torch.rand(8, 2, device=dev) * 3,
                   torch.rand(8, device=dev))
        assert p.shape == (8, 17, 3) and b.shape == (8, 6) and p.dtype == torch.float32


@pytest.mark.parametrize("n,h,w,dark,shift_heatmap", [
    (4096, 64, 48, False, False),   # BASELINE config 2 size, 9 stages
    (4096, 64, 48, True, True),
    (2048, 96, 72, True, False),    # BASELINE config 3 size: 4 stages < 7 consumer warps
    (2048, 96, 72, False, True),
    (1000, 128, 96, False, False),  # 2 stages
])
def test_decode_large_batch_properties(cuda_device, n, h, w, dark, shift_heatmap):
    """Full BASELINE sizes (flip pair): compare against torch reductions
    (size-independent properties: argmax value/index, idempotence of a second call).
    The stage counts differ per map size, which exercises every pipeline depth."""
    dev = cuda_device
    k = 17
    g = torch.Generator(device=dev).manual_seed(0)
    hm = torch.rand(n, k, h, w, device=dev, generator=g)
    fl = torch.rand(n, k, h, w, device=dev, generator=g)
    fidx = torch.as_tensor(synth.flip_index(), device=dev)
    center = torch.rand(n, 2, device=dev) * 400
    scale = torch.rand(n, 2, device=dev) * 2.8 + 0.2
    score = torch.rand(n, device=dev)
    dec = mp.create_decoder("topdown_heatmap", to_original=False, dark_udp_refine=dark)
    p, b = dec.decode_flip_pair(hm, fl, synth.flip_index(), center, scale, score,
                                shift_heatmap=shift_heatmap)
    fb = fl[:, fidx].flip(-1)
    if shift_heatmap:
        fb = torch.cat([fb[..., :1], fb[..., :-1]], dim=-1)
    avg = (hm + fb) * 0.5
    vals, idx = avg.reshape(n, k, -1).max(dim=2)
    assert torch.equal(p[..., 2], vals)
    if not dark:
        assert torch.equal(p[..., 0], (idx % w).float())
        assert torch.equal(p[..., 1], (idx // w).float())
    p2, b2 = dec.decode_flip_pair(hm, fl, synth.flip_index(), center, scale, score,
                                  shift_heatmap=shift_heatmap)
    assert torch.equal(p, p2) and torch.equal(b, b2)
    # no-flip path at the same size
    p3, _ = mp.create_decoder("topdown_heatmap", to_original=False)(hm, center, scale, score)
    vals, idx = hm.reshape(n, k, -1).max(dim=2)
    assert torch.equal(p3[..., 2], vals)
    assert torch.equal(p3[..., 0], (idx % w).float()) and torch.equal(p3[..., 1], (idx // w).float())


@pytest.mark.parametrize("n,h,w,kw,shift_heatmap", [
    (4096, 64, 48, dict(dark_udp_refine=True, kernel_size=11), False),   # bench.py headline
    (4096, 64, 48, dict(shift_coordinate=True), True),                   # BASELINE config 1 recipe
    (2048, 96, 72, dict(use_udp=True, dark_udp_refine=True, kernel_size=11), False),  # config 3
])
def test_decode_bench_sizes_match_oracle_on_a_subset(cuda_device, n, h, w, kw, shift_heatmap):
    """The bench's own workload (its blob generator, its batch sizes, its decoder settings, so
    its pipeline depth and consumer-group split): 256 crops spread over the whole batch are
    compared with the oracle -- maxvals and boxes bit for bit, refined coordinates to 1e-4 px
    (the unrefined / quarter-shift ones exactly), in heat-map and in image coordinates."""
    import bench

    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(21)
    fidx = synth.flip_index()
    hm, fl = bench.blob_stack(torch, n, h, w, dev, g, fidx)
    center = torch.rand(n, 2, device=dev, generator=g) * 400
    scale = torch.rand(n, 2, device=dev, generator=g) * 2.8 + 0.2
    score = torch.rand(n, device=dev, generator=g)
    pick = np.unique(np.concatenate([np.random.RandomState(3).choice(n, 250, replace=False),
                                     [0, 1, n // 2, n - 2, n - 1, 147, 148, 149]]))
    okw = dict(shift_coordinate_flag=kw.get("shift_coordinate", False),
               use_udp=kw.get("use_udp", False),
               dark_udp_refine_flag=kw.get("dark_udp_refine", False),
               kernel_size=kw.get("kernel_size", 11))
    refined = okw["dark_udp_refine_flag"]
    for to_original in (False, True):
        dec = mp.create_decoder("topdown_heatmap", to_original=to_original, **kw)
        p, b = dec.decode_flip_pair(hm, fl, fidx, center, scale, score, shift_heatmap=shift_heatmap)
        want_p, want_b = topdown_decode.decode_with_flip(
            hm[pick].cpu().numpy(), fl[pick].cpu().numpy(), fidx, center[pick].cpu().numpy(),
            scale[pick].cpu().numpy(), score[pick].cpu().numpy(), shift_heatmap=shift_heatmap,
            to_original=to_original, **okw)
        got_p, got_b = p[pick].cpu().numpy(), b[pick].cpu().numpy()
        assert np.array_equal(got_p[..., 2], want_p[..., 2])
        assert np.array_equal(got_b, want_b)
        if not refined:
            assert np.array_equal(got_p, want_p)
        else:
            # image coordinates = heat-map coordinates * scale * 200 / size: the tolerance scales
            tol = COORD_TOL if not to_original else \
                COORD_TOL * float(scale.max()) * 200.0 / min(h, w)
            assert np.abs(got_p[..., :2] - want_p[..., :2]).max() <= tol


def test_decode_subnormal_and_zero_maps(cuda_device):
    """Flip averaging of values whose halves are subnormal (the kernel compares sums and
    must fall back to exact averages there), all-zero and all-negative planes."""
    dev = cuda_device
    n, k, h, w = 4, 17, 64, 48
    rng = np.random.RandomState(7)
    tiny = np.float32(2.0 ** -149)
    maps = (rng.randint(0, 8, (n, k, h, w)).astype(np.float32) * tiny)
    flipped = (rng.randint(0, 8, (n, k, h, w)).astype(np.float32) * tiny)
    maps[1] = 0.0
    flipped[1] = 0.0
    maps[2] = -rng.random_sample((k, h, w)).astype(np.float32) - 1.0
    center, scale, score = synth.crop_geometry(n, seed=1)
    want_p, want_b = topdown_decode.decode_with_flip(maps, flipped, synth.flip_index(), center,
                                                     scale, score, to_original=False)
    dec = mp.create_decoder("topdown_heatmap", to_original=False)
    got_p, got_b = dec.decode_flip_pair(_t(maps, dev), _t(flipped, dev), synth.flip_index(),
                                        _t(center, dev), _t(scale, dev), _t(score, dev))
    assert np.array_equal(got_p.cpu().numpy(), want_p)
    assert np.array_equal(got_b.cpu().numpy(), want_b)


def test_decode_host_front_end_matches_device_path(cuda_device):
    n, h, w = 300, 64, 48
    maps, _ = synth.blob_heatmaps(n, 17, h, w, seed=2)
    flipped = synth.flipped_pair(maps, seed=2)
    center, scale, score = synth.crop_geometry(n, seed=2)
    dec = mp.create_decoder("topdown_heatmap", dark_udp_refine=True)
    p = dec._params(17, h, w, flip_index=synth.flip_index(), shift_heatmap=False)
    ctx = codec.HostContext(0, scratch_bytes=64 << 20)
    hp, hb = ctx.topdown_decode(maps, center, scale, score, flipped=flipped, params=p)
    ctx.close()
    dp, db = dec.decode_flip_pair(_t(maps, cuda_device), _t(flipped, cuda_device),
                                  synth.flip_index(), _t(center, cuda_device),
                                  _t(scale, cuda_device), _t(score, cuda_device))
    assert np.array_equal(hp, dp.cpu().numpy()) and np.array_equal(hb, db.cpu().numpy())



@pytest.mark.parametrize("use_udp", [False, True])
def test_affine_host_front_end_upload_modes_match_device_path(cuda_device, use_udp):
    """pc_topdown_affine_host with whole-image upload and with the two rectangle uploads
    gives the crops of the device path bit for bit -- with rotations, boxes that leave the
    image, a degenerate box, and the context scratch holding OTHER images first (a kernel
    that sampled outside the uploaded rectangle would read those)."""
    n, hs, ws = 40, 240, 320
    images, boxes = synth.source_images_and_boxes(n, hs, ws, seed=11)
    rng = np.random.RandomState(11)
    boxes[1] = [-60.0, -40.0, 150.0, 200.0]          # leaves the image top-left
    boxes[2] = [250.0, 180.0, 200.0, 160.0]          # leaves it bottom-right
    boxes[3] = [-500.0, -500.0, 50.0, 60.0]          # never touches it: all border
    boxes[4] = [100.0, 100.0, 0.0, 0.0]              # degenerate
    boxes[5] = [0.0, 0.0, float(ws), float(hs)]      # the whole image (+ padding)
    rot = np.zeros(n, np.float32)
    rot[::3] = rng.uniform(-80, 80, len(rot[::3])).astype(np.float32)
    rot[7] = 180.0
    cfg = synth.TOPDOWN_CONFIG
    dev = cuda_device
    bt = mp.create_transform("topdown_box_to_center_scale", is_train=False, config=cfg)
    at = mp.create_transform("topdown_affine", is_train=False, config=cfg, use_udp=use_udp)
    c, s = bt.box_to_center_scale_batch(_t(boxes, dev))
    want, _ = at.affine_batch(_t(images, dev), c, s, _t(rot, dev))
    want = want.cpu().numpy()
    other = rng.randint(0, 256, size=images.shape, dtype=np.uint8)
    ctx = codec.HostContext(0, scratch_bytes=32 << 20)    # several chunks per slot
    try:
        sent = {}
        pinned = torch.from_numpy(images).pin_memory()
        for mode, src in (("full", images), ("roi", images), ("roi_kernel", pinned.numpy()),
                          ("roi_kernel", images)):      # pageable: falls back to "roi"
            ctx.topdown_affine(other, boxes, cfg["image_size"], rot=rot, use_udp=use_udp,
                               upload="full")             # poison the scratch
            crops, hc, hsc = ctx.topdown_affine(src, boxes, cfg["image_size"], rot=rot,
                                                use_udp=use_udp, upload=mode)
            key = mode + ("_pageable" if mode == "roi_kernel" and src is images else "")
            sent[key] = ctx.last_transfer_bytes()
            bad = [i for i in range(n) if not np.array_equal(crops[i], want[i])]
            assert not bad, (key, bad)
            assert np.array_equal(hc, c.cpu().numpy()) and np.array_equal(hsc, s.cpu().numpy())
        assert sent["full"][0] >= images.size
        assert sent["roi"][0] < sent["roi_kernel"][0] < sent["full"][0]   # 64-byte row spans
        assert sent["roi_kernel_pageable"][0] == sent["roi"][0]
        with pytest.raises(ValueError):
            ctx.topdown_affine(images, boxes, cfg["image_size"], upload="some")
    finally:
        ctx.close()


# --------------------------------------------------------------------- encode
@pytest.mark.parametrize("cfg", [synth.TOPDOWN_CONFIG, synth.TOPDOWN_CONFIG_384])
@pytest.mark.parametrize("use_udp", [False, True])
def test_encode_matches_oracle(cuda_device, cfg, use_udp):
    n = 64
    kps = synth.keypoints(n, 17, cfg["image_size"], seed=0)
    t = mp.create_transform("topdown_generate_target", is_train=True, config=cfg, sigma=2.0,
                            use_udp=use_udp)
    target, weight = t.encode_batch(_t(kps, cuda_device))
    target, weight = target.cpu().numpy(), weight.cpu().numpy()
    fn = topdown_encode.encode_udp if use_udp else topdown_encode.encode_gaussian
    for i in range(n):
        want_t, want_w = fn(kps[i], cfg["image_size"], cfg["heatmap_size"], sigma=2.0)
        assert np.array_equal(weight[i], want_w)
        assert np.array_equal(target[i] != 0, want_t != 0)  # same support
        assert np.abs(target[i] - want_t).max() <= TARGET_TOL


def test_encode_matches_reference_golden(cuda_device, golden):
    g = golden("encode_ref.npz")
    for tag, cfg in (("64x48", synth.TOPDOWN_CONFIG), ("96x72", synth.TOPDOWN_CONFIG_384)):
        kps = g[f"kps_{tag}"]
        for name, udp in (("std", False), ("udp", True)):
            t = mp.create_transform("topdown_generate_target", config=cfg, use_udp=udp)
            target, weight = t.encode_batch(_t(kps, cuda_device))
            assert np.array_equal(weight.cpu().numpy(), g[f"weight_{name}_{tag}"])
            assert np.abs(target.cpu().numpy() - g[f"target_{name}_{tag}"]).max() <= TARGET_TOL
    cfg = dict(synth.TOPDOWN_CONFIG, joint_weights=g["joint_weights"].tolist())
    t = mp.create_transform("topdown_generate_target", config=cfg, sigma=3.0,
                            use_different_joint_weights=True)
    target, weight = t.encode_batch(_t(g["kps_64x48"][:4], cuda_device))
    assert np.abs(target.cpu().numpy() - g["target_std_sigma3"]).max() <= TARGET_TOL
    assert np.abs(weight.cpu().numpy() - g["weight_std_sigma3"]).max() <= 1e-6


def test_encode_per_sample_call_convention(cuda_device):
    """t(*columns) with the train column order, as MindSpore's dataset.map calls it."""
    cfg = synth.TOPDOWN_CONFIG
    t = mp.create_transform("topdown_generate_target", is_train=True, config=cfg)
    kps = synth.keypoints(1, 17, cfg["image_size"], seed=4)[0]
    cols = [np.zeros((256, 192, 3), np.uint8), np.zeros(2, np.float32), np.ones(2, np.float32),
            np.zeros(4, np.float32), kps, np.float32(0), np.zeros(1, np.float32),
            np.zeros(1, np.float32)]
    out = t(*cols)
    want_t, want_w = topdown_encode.encode_gaussian(kps, cfg["image_size"], cfg["heatmap_size"])
    assert len(out) == 8 and out[6].shape == (17, 64, 48)
    assert np.abs(out[6] - want_t).max() <= TARGET_TOL and np.array_equal(out[7], want_w)


def test_encode_decode_round_trip_full_size(cuda_device):
    """Property at BASELINE size: decoding an encoded in-bounds keypoint returns its pixel."""
    n = 2048
    cfg = synth.TOPDOWN_CONFIG
    rng = np.random.RandomState(0)
    kps = np.ones((n, 17, 3), np.float32)
    kps[..., 0] = rng.uniform(8, 180, (n, 17))
    kps[..., 1] = rng.uniform(8, 240, (n, 17))
    t = mp.create_transform("topdown_generate_target", config=cfg)
    target, weight = t.encode_batch(_t(kps, cuda_device))
    dec = mp.create_decoder("topdown_heatmap", to_original=False)
    z = torch.zeros(n, 2, device=cuda_device)
    p, _ = dec(target, z, z + 1, torch.zeros(n, device=cuda_device))
    mu = torch.from_numpy(np.rint(kps[..., :2].astype(np.float64) / 4.0).astype(np.float32))
    assert torch.equal(p[..., :2].cpu(), mu) and torch.all(p[..., 2] == 1.0)
    assert torch.all(weight == 1.0)


# ----------------------------------------------------------------- warp / geometry
def test_geometry_matches_oracle(cuda_device, golden):
    g = golden("affine_ref.npz")
    dev = cuda_device
    for tag, image_size in (("256x192", [192, 256]), ("384x288", [288, 384])):
        cfg = dict(synth.TOPDOWN_CONFIG, image_size=image_size)
        t = mp.create_transform("topdown_box_to_center_scale", is_train=False, config=cfg)
        c, s = t.box_to_center_scale_batch(_t(g["boxes"], dev))
        assert np.array_equal(c.cpu().numpy(), g[f"center_{tag}"])
        assert np.array_equal(s.cpu().numpy(), g[f"scale_{tag}"])
        rot = _t(g["rots"].astype(np.float32), dev)
        ref_std = np.stack([affine.affine_matrix(ci, si, float(np.float32(r)), np.array(image_size))
                            for ci, si, r in zip(g[f"center_{tag}"], g[f"scale_{tag}"], g["rots"])])
        fwd, inv = codec.affine_matrices(c, s, rot, image_size, use_udp=False)
        zero_rot = g["rots"] == 0
        # rot = 0 (the eval path): bit-identical to cv2.getAffineTransform; rot != 0 goes
        # through the device's sin/cos (<= 1 ulp from numpy's before the float32 point store)
        assert np.array_equal(fwd.cpu().numpy()[zero_rot], g[f"std_{tag}"][zero_rot])
        assert np.allclose(fwd.cpu().numpy(), ref_std, rtol=0, atol=1e-9 * np.abs(ref_std).max())
        want_inv = np.stack([warp.invert_affine(m) for m in fwd.cpu().numpy()])
        assert np.array_equal(inv.cpu().numpy(), want_inv)
        ref_udp = np.stack([affine.udp_matrix(ci, si, float(np.float32(r)), np.array(image_size))
                            for ci, si, r in zip(g[f"center_{tag}"], g[f"scale_{tag}"], g["rots"])])
        fwd_u, _ = codec.affine_matrices(c, s, rot, image_size, use_udp=True)
        fu = fwd_u.cpu().numpy()
        assert np.array_equal(fu[zero_rot], ref_udp[zero_rot].astype(np.float64))
        assert np.allclose(fu, ref_udp, rtol=2e-7, atol=1e-6)
        # joints through the reference's own matrices
        k_std = codec.affine_joints(_t(g["kps_in"], dev).clone(), _t(g[f"std_{tag}"], dev), False)
        assert np.abs(k_std.cpu().numpy() - g[f"kps_std_{tag}"]).max() == 0
        k_udp = codec.affine_joints(_t(g["kps_in"], dev).clone(),
                                    _t(g[f"udp_{tag}"].astype(np.float64), dev), True)
        # sgemm accumulation order is the BLAS build's; 2 ulp is the bar, 0 is what we see
        assert np.allclose(k_udp.cpu().numpy(), g[f"kps_udp_{tag}"], rtol=3e-7, atol=1e-5)


@pytest.mark.parametrize("use_udp", [False, True])
def test_rotated_per_sample_transform_matches_reference_golden(cuda_device, golden, use_udp):
    """(center, scale, rot) -> crop for rot != 0, end to end against the reference's OWN
    TopDownAffine.transform (its matrix, cv2.warpAffine, its joints) fed float32 rotations as
    its pipeline does: the per-sample transform evaluates the rotation scalars with numpy on
    the host and solves / warps on the device -- every crop byte equal.  The batched device
    path (device sin / cos, matrix within 1e-9) is measured beside it."""
    g = golden("affine_rot_ref.npz")
    tag = "udp" if use_udp else "std"
    cfg = dict(synth.TOPDOWN_CONFIG, image_size=[96, 128], heatmap_size=[24, 32])
    at = mp.create_transform("topdown_affine", is_train=False, config=cfg, use_udp=use_udp)
    n = len(g["rots"])
    for i in range(n):
        state = dict(image=g["images"][i], center=g["center"][i], scale=g["scale"][i],
                     rotation=np.asarray(g["rots"][i]), keypoints=g["kps_in"][i].copy())
        out = at.transform(state)
        assert np.array_equal(out["image"], g[f"crops_{tag}"][i]), i
        if use_udp:   # sgemm accumulation order is the BLAS build's (2 ulp bar, as at rot = 0)
            assert np.allclose(out["keypoints"], g[f"kps_{tag}"][i], rtol=3e-7, atol=1e-5)
        else:
            assert np.array_equal(out["keypoints"], g[f"kps_{tag}"][i])
    # batched path, rotation on the device: how many bytes differ from the reference's crops
    dev = cuda_device
    crops, _ = at.affine_batch(_t(g["images"], dev), _t(g["center"], dev), _t(g["scale"], dev),
                               _t(g["rots"], dev))
    diff = int((crops.cpu().numpy() != g[f"crops_{tag}"]).sum())
    print(f"\nbatched rotated crops ({tag}): {diff} of {crops.numel()} bytes differ from the "
          "reference's (device sin/cos, float64 direction)")
    assert diff <= crops.numel() // 50   # a matrix a few 1e-9 off moves a few 1/1024-pixel ties


def test_warp_matches_cv2_golden(cuda_device, golden):
    g = golden("warp_ref.npz")
    src, dst, sizes, mats = g["src"], g["dst"], g["sizes"], g["mats"]
    dev = cuda_device
    so = do = 0
    for (hs, ws, dw, dh), m in zip(sizes, mats):
        s = src[so:so + hs * ws * 3].reshape(1, hs, ws, 3)
        d = dst[do:do + dh * dw * 3].reshape(dh, dw, 3)
        so += hs * ws * 3
        do += dh * dw * 3
        inv = codec.invert_affine(_t(m[None], dev))
        out = codec.warp_affine_uniform(_t(s, dev), inv, (int(dw), int(dh)))
        assert np.array_equal(out[0].cpu().numpy(), d)


@pytest.mark.parametrize("use_udp", [False, True])
def test_warp_batch_matches_oracle(cuda_device, use_udp):
    n = 24
    images, boxes = synth.source_images_and_boxes(n, 240, 320, seed=1)
    cfg = synth.TOPDOWN_CONFIG
    dev = cuda_device
    bt = mp.create_transform("topdown_box_to_center_scale", is_train=False, config=cfg)
    at = mp.create_transform("topdown_affine", is_train=False, config=cfg, use_udp=use_udp)
    c, s = bt.box_to_center_scale_batch(_t(boxes, dev))
    rot = torch.zeros(n, device=dev)
    rot[::3] = 30.0
    kps = synth.keypoints(n, 17, [320, 240], seed=2)
    crops, kout = at.affine_batch(_t(images, dev), c, s, rot, _t(kps, dev))
    fwd, _ = codec.affine_matrices(c, s, rot, cfg["image_size"], use_udp=use_udp)
    fwd = fwd.cpu().numpy()
    crops = crops.cpu().numpy()
    for i in range(n):
        want = warp.warp_affine_u8(images[i], fwd[i], (192, 256))
        assert np.array_equal(crops[i], want), i
    try:
        import cv2
    except ImportError:
        cv2 = None
    if cv2 is not None:
        for i in range(0, n, 5):
            m = fwd[i].astype(np.float32) if use_udp else fwd[i]
            assert np.array_equal(crops[i], cv2.warpAffine(images[i], m, (192, 256),
                                                           flags=cv2.INTER_LINEAR))
    assert kout.shape == (n, 17, 3)


def test_affine_per_sample_call_convention(cuda_device):
    cfg = synth.TOPDOWN_CONFIG
    images, boxes = synth.source_images_and_boxes(1, 200, 300, seed=3)
    bt = mp.create_transform("topdown_box_to_center_scale", is_train=False, config=cfg)
    at = mp.create_transform("topdown_affine", is_train=False, config=cfg)
    cols = [images[0], np.zeros(2, np.float32), np.ones(2, np.float32), np.float32(0),
            np.array("x.jpg"), boxes[0], np.int32(0), np.float32(1)]
    cols = list(bt(*cols))
    c_want, s_want = affine.box_to_center_scale(tuple(boxes[0]), np.array(cfg["image_size"]))
    assert np.array_equal(cols[1], c_want) and np.array_equal(cols[2], s_want)
    out = at(*cols)
    m = affine.affine_matrix(c_want, s_want, 0.0, np.array(cfg["image_size"]))
    assert np.array_equal(out[0], warp.warp_affine_u8(images[0], m, (192, 256)))


@pytest.mark.parametrize("channels", [1, 3, 4])
def test_warp_image_corners_and_unaligned_sources(cuda_device, channels):
    """Identity-like matrices sample the first and last pixel pair of the source (the
    3-channel fast path reads aligned words and must not leave the image), and source
    images that start at odd byte offsets inside a shared buffer."""
    dev = cuda_device
    rng = np.random.RandomState(11)
    hs, ws, dw, dh = 37, 53, 48, 64
    mats = np.array([[[1, 0, 0], [0, 1, 0]],
                     [[1, 0, 0.5], [0, 1, 0.5]],
                     [[1, 0, -3.25], [0, 1, 2.75]],
                     [[0.9, 0.1, 1.0], [-0.1, 0.9, 2.0]],
                     [[1.7, 0, -20.0], [0, 1.7, -10.0]]], np.float64)
    n = len(mats)
    pad = [0, 1, 2, 3, 5]
    buf = rng.randint(0, 256, size=n * (hs * ws * channels + 8), dtype=np.uint8)
    offs, imgs, pos = [], [], 0
    for i in range(n):
        pos += pad[i]
        offs.append(pos)
        imgs.append(buf[pos:pos + hs * ws * channels].reshape(hs, ws, channels))
        pos += hs * ws * channels
    inv = codec.invert_affine(_t(mats, dev))
    out = codec.warp_affine(_t(buf, dev), torch.tensor(offs, device=dev),
                            torch.tensor([[hs, ws]] * n, device=dev, dtype=torch.int32), inv,
                            (dw, dh), channels=channels)
    out = out.cpu().numpy()
    for i in range(n):
        want = warp.warp_affine_u8(imgs[i], mats[i], (dw, dh))
        assert np.array_equal(out[i], want.reshape(dh, dw, channels)), i


@pytest.mark.parametrize("ws,dw,dh", [(52, 64, 40), (64, 32, 16), (40, 96, 50), (320, 192, 256)])
def test_warp_band_path_edge_cases(cuda_device, ws, dw, dh):
    """The shared-memory band kernel (3 channels, dst_w % 32 == 0, no rotation, ws % 4 == 0):
    source rows whose 16-byte phase changes from row to row (ws * 3 % 16 != 0), images at odd
    byte offsets, crops that hang over every edge or miss the image, up- and down-scaling up
    to several band passes per tile, and the matrices that must take the quad path instead
    (mirrored, slightly rotated, a band too large for shared memory)."""
    dev = cuda_device
    rng = np.random.RandomState(ws + dw)
    hs = 37
    mats = np.array([
        [[1, 0, 0], [0, 1, 0]],                    # identity: first / last pixel pair
        [[1, 0, 0.5], [0, 1, 0.5]],
        [[1, 0, -3.25], [0, 1, 2.75]],             # over the left / bottom edge
        [[1, 0, 7.0], [0, 1, -5.5]],               # over the right / top edge
        [[3.7, 0, -11.0], [0, 3.7, -9.0]],         # up-scaling: rows and columns repeat
        [[0.6, 0, 3.0], [0, 0.6, 2.0]],            # down-scaling by 1.67
        [[0.21, 0, 1.0], [0, 0.23, 0.5]],          # ... by 4.8 / 4.3 (anisotropic)
        [[0.02, 0, 0.0], [0, 0.02, 0.0]],          # ... by 50: larger than the image
        [[1, 0, 500.0], [0, 1, 0]],                # misses the image
        [[1, 0, 0], [0, 1, -300.0]],
        [[-1, 0, 40.0], [0, 1, 0]],                # mirrored: quad path
        [[1, 0.002, 0], [-0.002, 1, 0]],           # a whisper of rotation: quad path
        [[1.3, 0, -1e-7], [0, 1.3, 1e-7]],
    ], np.float64)
    n = len(mats)
    pad = [(3 * i) % 7 for i in range(n)]          # odd byte offsets inside a shared buffer
    buf = rng.randint(0, 256, size=n * (hs * ws * 3 + 8), dtype=np.uint8)
    offs, imgs, pos = [], [], 0
    for i in range(n):
        pos += pad[i]
        offs.append(pos)
        imgs.append(buf[pos:pos + hs * ws * 3].reshape(hs, ws, 3))
        pos += hs * ws * 3
    inv = codec.invert_affine(_t(mats, dev))
    out = codec.warp_affine(_t(buf, dev), torch.tensor(offs, device=dev),
                            torch.tensor([[hs, ws]] * n, device=dev, dtype=torch.int32), inv,
                            (dw, dh), channels=3).cpu().numpy()
    for i in range(n):
        want = warp.warp_affine_u8(imgs[i], mats[i], (dw, dh))
        assert np.array_equal(out[i], want.reshape(dh, dw, 3)), (i, mats[i].tolist())


def test_warp_band_path_tiny_and_large_sources(cuda_device):
    """1- and 2-pixel images (both zero-fill ranges of a band row meet), and a 1080 x 1920
    source scaled down by 7.5 (one output row per band pass)."""
    dev = cuda_device
    rng = np.random.RandomState(5)
    for hs, ws, m in ((1, 4, [[1, 0, 1.0], [0, 1, 3.0]]), (2, 4, [[4, 0, 8.0], [0, 4, 2.0]]),
                      (1080, 1920, [[0.1333, 0, -3.0], [0, 0.1333, -4.0]]),
                      (1080, 1920, [[0.4, 0, -300.0], [0, 0.4, -100.0]])):
        img = rng.randint(0, 256, size=(1, hs, ws, 3), dtype=np.uint8)
        mat = np.array([m], np.float64)
        out = codec.warp_affine_uniform(_t(img, dev), codec.invert_affine(_t(mat, dev)), (256, 144))
        want = warp.warp_affine_u8(img[0], mat[0], (256, 144))
        assert np.array_equal(out[0].cpu().numpy(), want), (hs, ws)


def test_warp_fused_normalize_chw(cuda_device):
    """N2: warp + Normalize(mean * 255, std * 255) + HWC2CHW in one kernel equals the uint8
    warp followed by the oracle's normalisation, bit for bit (same float32 formula)."""
    n = 12
    images, boxes = synth.source_images_and_boxes(n, 240, 320, seed=3)
    cfg = synth.TOPDOWN_CONFIG
    dev = cuda_device
    bt = mp.create_transform("topdown_box_to_center_scale", is_train=False, config=cfg)
    at = mp.create_transform("topdown_affine", is_train=False, config=cfg)
    c, s = bt.box_to_center_scale_batch(_t(boxes, dev))
    rot = torch.zeros(n, device=dev)
    rot[::4] = -25.0
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.255]  # the reference's defaults
    crops_u8, _ = at.affine_batch(_t(images, dev), c, s, rot)
    fused, _ = at.affine_batch(_t(images, dev), c, s, rot, normalize_mean=mean, normalize_std=std)
    assert fused.shape == (n, 3, 256, 192) and fused.dtype == torch.float32
    m255 = (np.array(mean) * 255.0).tolist()
    s255 = (np.array(std) * 255.0).tolist()
    for i in range(n):
        want = warp.normalize_chw(crops_u8[i].cpu().numpy(), m255, s255)
        assert np.array_equal(fused[i].cpu().numpy(), want), i
    # other statistics (the library checks on the host, for all 256 pixel values of a channel,
    # whether its 3-instruction quotient equals the IEEE division, and divides otherwise)
    for mean2, std2 in (([0.3, 0.1, 0.7], [0.11, 0.37, 0.013]), ([0.0, 0.5, 1.0], [1.0, 0.333, 3.0])):
        m2, s2 = (np.array(mean2) * 255.0).tolist(), (np.array(std2) * 255.0).tolist()
        fused2, _ = at.affine_batch(_t(images, dev), c, s, rot, normalize_mean=mean2,
                                    normalize_std=std2)
        for i in range(n):
            want = warp.normalize_chw(crops_u8[i].cpu().numpy(), m2, s2)
            assert np.array_equal(fused2[i].cpu().numpy(), want), (mean2, i)
    with pytest.raises(ValueError):
        codec.warp_affine_normalized(_t(images, dev), torch.zeros(n, device=dev),
                                     torch.zeros(n, 2, device=dev), torch.zeros(n, 2, 3, device=dev),
                                     (190, 256), m255, s255)   # dst_w % 4 != 0
