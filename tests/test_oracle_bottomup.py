"""Bottom-up oracle: LSAP vs scipy goldens, match_by_tag vs reference goldens,
decode restatement frozen + properties."""
import numpy as np
import pytest

from mindpose_b200 import synth
from oracle import bottomup_decode as bd
from oracle import bottomup_encode as be
from oracle import gen_golden_bottomup as ggb
from oracle import grouping, lsap, ref_loader
from oracle import refine_missing as rm


def _split_counts(flat, counts, k=17, width=4):
    out, off = [], 0
    for c in counts:
        size = int(c) * k * width
        out.append(flat[off:off + size].reshape(int(c), k, width))
        off += size
    return out


def test_lsap_matches_scipy_golden(golden):
    g = golden("lsap_ref.npz")
    co = ro = 0
    for nr, nc in g["shapes"]:
        c = g["cost"][co:co + nr * nc].reshape(nr, nc)
        co += nr * nc
        m = min(nr, nc)
        want_r, want_c = g["rows"][ro:ro + m], g["cols"][ro:ro + m]
        ro += m
        r, cc = lsap.linear_sum_assignment(c)
        assert np.array_equal(r, want_r) and np.array_equal(cc, want_c)


def test_lsap_matches_installed_scipy():
    so = pytest.importorskip("scipy.optimize")
    rng = np.random.RandomState(9)
    for it in range(500):
        nr, nc = rng.randint(1, 10), rng.randint(1, 12)
        c = rng.randint(0, 3, (nr, nc)).astype(float) if it % 2 else np.round(rng.uniform(0, 4, (nr, nc)))
        r1, c1 = so.linear_sum_assignment(c)
        r2, c2 = lsap.linear_sum_assignment(c)
        assert np.array_equal(r1, r2) and np.array_equal(c1, c2)


@pytest.mark.parametrize("mode", ["people", "ties", "crowded"])
@pytest.mark.parametrize("rounded", [True, False])
def test_match_by_tag_matches_reference_golden(golden, mode, rounded):
    g = golden("match_ref.npz")
    name = f"{mode}_{'rounded' if rounded else 'exact'}"
    want = _split_counts(g[f"ans_{name}"], g[f"counts_{name}"])
    val, tag, ind = g[f"val_{mode}"], g[f"tag_{mode}"], g[f"ind_{mode}"]
    for i in range(val.shape[0]):
        got = grouping.match_by_tag(val[i], tag[i], ind[i], synth.COCO_JOINT_ORDER,
                                    vis_thr=0.1, tag_thr=1.0, use_rounded_norm=rounded)
        if want[i].shape[0] == 0:
            assert got.shape == (0,)
        else:
            assert np.array_equal(got, want[i]), (mode, rounded, i)


def test_match_by_tag_key_collision_resets_group():
    """Two detections opening a group with the same tag value share one group (dict key)."""
    val = np.zeros((17, 30), np.float32)
    tag = np.zeros((17, 30, 1), np.float32)
    ind = np.zeros((17, 30, 2), np.float32)
    val[0, :2] = [0.9, 0.8]
    tag[0, :2, 0] = [5.0, 5.0]
    ind[0, 0] = [1, 2]
    ind[0, 1] = [3, 4]
    got = grouping.match_by_tag(val, tag, ind, synth.COCO_JOINT_ORDER)
    assert got.shape == (1, 17, 4) and got[0, 0].tolist() == [3.0, 4.0, np.float32(0.8), 5.0]


def test_instance_score_is_numpy_mean():
    rng = np.random.RandomState(0)
    ans = rng.uniform(0, 1, (5, 17, 4)).astype(np.float32)
    want = [y[:, 2].mean() for y in ans]
    assert grouping.instance_scores(ans) == want


@pytest.mark.needs_reference
def test_match_by_tag_against_live_reference():
    from oracle.gen_golden_bottomup import grouping_inputs

    ns = ref_loader.load()
    for mode in ("people", "ties", "crowded"):
        val, tag, ind = grouping_inputs(77, 12, mode=mode)
        for i in range(val.shape[0]):
            want = ns.match.match_by_tag(val[i], tag[i], ind[i], synth.COCO_JOINT_ORDER)
            got = grouping.match_by_tag(val[i], tag[i], ind[i], synth.COCO_JOINT_ORDER)
            assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.needs_reference
def test_transform_keypoints_against_live_reference():
    ns = ref_loader.load()
    rng = np.random.RandomState(1)
    coords = [rng.uniform(0, 256, (3, 17, 4)).astype(np.float32), np.zeros((0,), np.float32)]
    center = rng.uniform(100, 300, (2, 2))
    scale = rng.uniform(1, 3, (2, 2))
    hw = np.array([[256.0, 256.0], [256.0, 192.0]])
    want = ns.utils.transform_keypoints(coords, center, scale, hw)
    got = grouping.transform_keypoints(coords, center, scale, hw)
    assert all(np.array_equal(a, b) for a, b in zip(got, want))


# ------------------------------------------------------------ decode restatement
def test_bottomup_decode_restatement_is_frozen(golden):
    g = golden("bottomup_decode_restated.npz")
    d = synth.bottomup_outputs(2, 17, 32, 32, mask_hw=(128, 128), seed=4, max_people=4)
    val_k, tag_k, ind_k, raw, tagging = bd.decode([d["out0"], d["out1"]], d["mask"],
                                                  use_nms=True, nms_kernel=3, max_num=30)
    assert np.array_equal(val_k, g["val_k"]) and np.array_equal(ind_k, g["ind_k"])
    assert np.array_equal(tag_k, g["tag_k"])
    assert val_k.shape == (2, 17, 30) and tag_k.shape == (2, 17, 30, 1) and ind_k.shape == (2, 17, 30, 2)
    assert raw.shape == (2, 17, 64, 64) and tagging.shape == (2, 17, 64, 64, 1)


def test_resize_bilinear_legacy_2x_is_average_of_neighbours():
    x = np.arange(12, dtype=np.float32).reshape(1, 1, 3, 4)
    y = bd.resize_bilinear_legacy(x, 6, 8)[0, 0]
    assert np.array_equal(y[::2, ::2], x[0, 0])
    assert np.array_equal(y[0, 1::2], [0.5, 1.5, 2.5, 3.0])   # last column clamps
    assert np.array_equal(y[5], y[4])                           # last row clamps


def test_topk_ties_take_lowest_index_and_nms_keeps_plateaus():
    heat = np.zeros((1, 1, 8, 8), np.float32)
    heat[0, 0, 2, 2] = heat[0, 0, 2, 3] = 0.5     # plateau: both survive a 3x3 NMS
    heat[0, 0, 6, 6] = 0.9
    kept = bd.nms(heat, 3)
    assert kept[0, 0, 2, 2] == 0.5 and kept[0, 0, 2, 3] == 0.5
    val, tag, ind, order = bd.top_k(kept, np.zeros((1, 1, 8, 8, 1), np.float32), 5)
    assert order[0, 0].tolist() == [54, 18, 19, 0, 1]
    assert val[0, 0].tolist() == [np.float32(0.9), 0.5, 0.5, 0.0, 0.0]


# ------------------------------------------------------------------ N1: bottom-up encode
@pytest.mark.parametrize("case", range(len(ggb.BOTTOMUP_ENCODE_CASES)))
def test_bottomup_encode_matches_reference_golden(golden, case):
    g = golden("bottomup_encode_ref.npz")
    sizes, m, tpj, seed = ggb.BOTTOMUP_ENCODE_CASES[case]
    kps = ggb.bottomup_people(seed, m, 17, sizes)
    target, tag_ind = be.encode(kps, sizes, sigma=2.0, max_num=30, tag_per_joint=tpj)
    assert target.dtype == np.float32 and tag_ind.dtype == np.int32
    assert np.array_equal(target, g[f"target_{case}"])
    assert np.array_equal(tag_ind, g[f"tag_ind_{case}"])


def test_bottomup_encode_rejects_too_many_people():
    kps = np.zeros((31, 17, 3), np.float32)
    with pytest.raises(ValueError, match="exeeds the maximum num"):
        be.generate_heatmap_and_tag_ind(kps, (64, 64), max_num=30)


@pytest.mark.needs_reference
def test_bottomup_encode_matches_live_reference():
    ns = ref_loader.load()
    sizes = [[40, 24], [80, 48]]
    cfg = dict(image_size=[512, 512], max_image_size=[832, 512], heatmap_sizes=sizes,
               flip_pairs=[[1, 2]], pixel_std=200.0, tag_per_joint=True)
    t = ns.bottomup.BottomUpGenerateTarget(is_train=True, config=cfg, sigma=2.0, max_num=30)
    for seed in range(5):
        kps = ggb.bottomup_people(100 + seed, 9, 17, sizes)
        want = t.transform(dict(keypoints=[k.copy() for k in kps]))
        target, tag_ind = be.encode(kps, sizes)
        assert np.array_equal(target, want["target"]) and np.array_equal(tag_ind, want["tag_ind"])


# ------------------------------------------------------------------ N3: refine_missing
@pytest.mark.parametrize("seed", range(4))
def test_refine_missing_matches_reference_golden(golden, seed):
    g = golden("refine_missing_ref.npz")
    heat, tagm, kps = g[f"heat_{seed}"], g[f"tag_{seed}"], g[f"kps_{seed}"]
    got = np.stack([rm.refine_missing(heat, tagm, kp) for kp in kps])
    assert np.array_equal(got, g[f"refined_{seed}"])
    assert not np.array_equal(got, kps)  # something was filled in


@pytest.mark.needs_reference
def test_refine_missing_matches_live_reference_body():
    f = rm.reference_function()
    for seed in range(10, 14):
        heat, tagm, kps = ggb.refine_inputs(seed, h=24, w=40, people=4)
        for kp in kps:
            assert np.array_equal(rm.refine_missing(heat, tagm, kp), f(None, heat, tagm, kp.copy()))
