"""The oracle against the reference's own outputs (tests/golden/*_ref.npz) and,
in the build container, against the live reference / cv2."""
import numpy as np
import pytest

from mindpose_b200 import synth
from oracle import affine, ref_loader, topdown_decode, topdown_encode, warp


# ------------------------------------------------------------------ geometry
@pytest.mark.parametrize("tag,image_size", [("256x192", [192, 256]), ("384x288", [288, 384])])
def test_box_center_scale_and_matrices_match_reference(golden, tag, image_size):
    g = golden("affine_ref.npz")
    boxes, rots = g["boxes"], g["rots"]
    for i, b in enumerate(boxes):
        c, s = affine.box_to_center_scale(tuple(b), np.array(image_size))
        assert np.array_equal(c, g[f"center_{tag}"][i])
        assert np.array_equal(s, g[f"scale_{tag}"][i])
        m = affine.affine_matrix(c, s, float(rots[i]), np.array(image_size))
        ref = g[f"std_{tag}"][i]
        # cv2.getAffineTransform's LU is restated op for op: bit-identical
        assert np.array_equal(m, ref), i
        u = affine.udp_matrix(c, s, float(rots[i]), np.array(image_size))
        assert u.dtype == np.float32 and np.array_equal(u, g[f"udp_{tag}"][i])
        k1 = affine.transform_joints(g["kps_in"][i], ref)
        assert np.array_equal(k1, g[f"kps_std_{tag}"][i])
        k2 = affine.transform_joints_udp(g["kps_in"][i], g[f"udp_{tag}"][i])
        assert np.array_equal(k2, g[f"kps_udp_{tag}"][i])


# -------------------------------------------------------------------- encode
@pytest.mark.parametrize("tag,image_size,heatmap_size",
                         [("64x48", [192, 256], [48, 64]), ("96x72", [288, 384], [72, 96])])
def test_encode_matches_reference(golden, tag, image_size, heatmap_size):
    g = golden("encode_ref.npz")
    kps = g[f"kps_{tag}"]
    for name, fn in (("std", topdown_encode.encode_gaussian), ("udp", topdown_encode.encode_udp)):
        for i, k in enumerate(kps):
            t, w = fn(k, image_size, heatmap_size, sigma=2.0)
            assert t.dtype == np.float32
            assert np.array_equal(t, g[f"target_{name}_{tag}"][i]), (name, i)
            assert np.array_equal(w, g[f"weight_{name}_{tag}"][i]), (name, i)


def test_encode_sigma3_joint_weights(golden):
    g = golden("encode_ref.npz")
    for i, k in enumerate(g["kps_64x48"][:4]):
        t, w = topdown_encode.encode_gaussian(k, [192, 256], [48, 64], sigma=3.0,
                                              joint_weights=g["joint_weights"])
        assert np.array_equal(t, g["target_std_sigma3"][i])
        assert np.array_equal(w, g["weight_std_sigma3"][i])


# ---------------------------------------------------------------------- warp
def _warp_cases(g):
    src, dst, sizes, mats = g["src"], g["dst"], g["sizes"], g["mats"]
    so = do = 0
    for (hs, ws, dw, dh), m in zip(sizes, mats):
        s = src[so:so + hs * ws * 3].reshape(hs, ws, 3)
        d = dst[do:do + dh * dw * 3].reshape(dh, dw, 3)
        so += hs * ws * 3
        do += dh * dw * 3
        yield s, m, (int(dw), int(dh)), d


def test_warp_matches_cv2_golden(golden):
    for s, m, dsize, d in _warp_cases(golden("warp_ref.npz")):
        out = warp.warp_affine_u8(s, m, dsize)
        assert np.array_equal(out, d)


def test_warp_weight_table_is_closed_form():
    tab = warp.bilinear_weight_table().reshape(32, 32, 4)
    fy, fx = np.mgrid[0:32, 0:32]
    want = np.stack([(32 - fx) * (32 - fy), fx * (32 - fy), (32 - fx) * fy, fx * fy], -1) * 32
    assert np.array_equal(tab, want)


def test_warp_matches_installed_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(5)
    for i in range(12):
        img = rng.randint(0, 256, (rng.randint(40, 200), rng.randint(40, 260), 3)).astype(np.uint8)
        c = np.array([rng.uniform(0, img.shape[1]), rng.uniform(0, img.shape[0])], np.float32)
        s = np.array([rng.uniform(0.2, 1.5), rng.uniform(0.2, 1.5)], np.float32)
        m = affine.affine_matrix(c, s, float(rng.uniform(-45, 45)), np.array([96, 128]))
        ref = cv2.warpAffine(img, m, (96, 128), flags=cv2.INTER_LINEAR)
        assert np.array_equal(warp.warp_affine_u8(img, m, (96, 128)), ref)


# --------------------------------------------------- decode (restated, frozen)
@pytest.mark.parametrize("tag,h,w", [("64x48", 64, 48), ("96x72", 96, 72)])
def test_decode_restatement_is_frozen(golden, tag, h, w):
    g = golden("topdown_decode_restated.npz")
    fidx = synth.flip_index()
    n = 6
    blobs, centres = synth.blob_heatmaps(n, 17, h, w, seed=1)
    flipped = synth.flipped_pair(blobs, seed=1)
    center, scale, score = g[f"center_{tag}"], g[f"scale_{tag}"], g[f"score_{tag}"]
    p, b = topdown_decode.decode(blobs, center, scale, score)
    assert np.array_equal(p, g[f"plain_preds_{tag}"]) and np.array_equal(b, g[f"plain_boxes_{tag}"])
    p, _ = topdown_decode.decode(blobs, center, scale, score, shift_coordinate_flag=True)
    assert np.array_equal(p, g[f"shift_preds_{tag}"])
    p, _ = topdown_decode.decode(blobs, center, scale, score, dark_udp_refine_flag=True, use_udp=True)
    assert np.allclose(p, g[f"dark_udp_preds_{tag}"], rtol=0, atol=1e-4)
    p, _ = topdown_decode.decode_with_flip(blobs, flipped, fidx, center, scale, score,
                                           shift_heatmap=True, shift_coordinate_flag=True)
    assert np.array_equal(p, g[f"flip_shift_preds_{tag}"])
    p, _ = topdown_decode.decode_with_flip(blobs, flipped, fidx, center, scale, score,
                                           dark_udp_refine_flag=True)
    assert np.allclose(p, g[f"flip_dark_preds_{tag}"], rtol=0, atol=1e-4)


def test_decode_argmax_is_first_occurrence():
    hm = np.zeros((1, 2, 8, 12), np.float32)
    hm[0, 0, 3, 5] = hm[0, 0, 6, 1] = 2.0
    hm[0, 1, 7, 11] = 1.0
    coords, maxvals, idx = topdown_decode.max_preds(hm)
    assert idx.tolist() == [[3 * 12 + 5, 7 * 12 + 11]]
    assert coords[0, 0].tolist() == [5.0, 3.0] and maxvals[0, 0, 0] == 2.0


def test_dark_recovers_subpixel_centre():
    """Property of the DARK restatement: on clean sigma=2 blobs it finds the true centre."""
    maps, centres = synth.blob_heatmaps(4, 17, 64, 48, seed=9, noise=0.0)
    coords, _, _ = topdown_decode.max_preds(maps)
    ref = topdown_decode.dark_udp_refine(coords, maps, topdown_decode.dark_gaussian_kernel(11))
    # zero 'same' padding truncates the blur within 6 px of the border: interior blobs only
    cx, cy = centres[..., 0], centres[..., 1]
    inner = (cx > 8) & (cx < 48 - 9) & (cy > 8) & (cy < 64 - 9)
    assert inner.sum() > 20
    assert np.abs(ref - centres)[inner].max() < 1e-3


def test_decoder_rejects_dark_with_shift():
    with pytest.raises(ValueError):
        topdown_decode.decode(np.zeros((1, 1, 4, 4), np.float32), np.zeros((1, 2)),
                              np.ones((1, 2)), np.zeros(1), shift_coordinate_flag=True,
                              dark_udp_refine_flag=True)


# ------------------------------------------------- live reference (container)
@pytest.mark.needs_reference
def test_encode_against_live_reference():
    ns = ref_loader.load()
    rng = np.random.RandomState(123)
    cfg = dict(synth.TOPDOWN_CONFIG)
    for udp in (False, True):
        t = ns.topdown.TopDownGenerateTarget(is_train=True, config=cfg, use_udp=udp)
        fn = topdown_encode.encode_udp if udp else topdown_encode.encode_gaussian
        for _ in range(20):
            kp = synth.keypoints(1, 17, cfg["image_size"], seed=rng.randint(1 << 30))[0]
            ref = t.transform(dict(keypoints=kp.copy()))
            got_t, got_w = fn(kp, cfg["image_size"], cfg["heatmap_size"])
            assert np.array_equal(got_t, ref["target"])
            assert np.array_equal(got_w, ref["target_weight"])


def test_host_rotation_geometry_reproduces_the_reference_matrices(golden):
    """The few rotation scalars the per-sample TopDownAffine evaluates with numpy on the host
    (float32 rotation in, as from the reference's dataset pipeline) give the reference's own
    matrices bit for bit: UDP directly, the standard one through cv2.getAffineTransform of the
    three points (the device solve is that routine op for op)."""
    cv2 = pytest.importorskip("cv2")
    import mindpose_b200 as mp
    from mindpose_b200 import synth

    g = golden("affine_rot_ref.npz")
    cfg = dict(synth.TOPDOWN_CONFIG, image_size=[96, 128], heatmap_size=[24, 32])
    for tag, udp in (("std", False), ("udp", True)):
        at = mp.create_transform("topdown_affine", is_train=False, config=cfg, use_udp=udp)
        for i in range(len(g["rots"])):
            geo = at._host_geometry(g["center"][i], g["scale"][i], np.asarray(g["rots"][i]))
            m = geo[1].astype(np.float64) if geo[0] == "matrix" else \
                cv2.getAffineTransform(geo[1], geo[2])
            assert np.array_equal(m, g[f"mats_{tag}"][i]), (tag, i)
