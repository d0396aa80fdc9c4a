"""Row N4 oracle (oracle/nms.py) against vectors produced by the unmodified reference
(mindpose/utils/nms.py and TopDownEvaluator.eval; see oracle/gen_golden_nms.py)."""
import numpy as np
import pytest

from oracle import gen_golden_nms as ggn
from oracle import nms, ref_loader


@pytest.mark.parametrize("ci", range(len(ggn.NMS_CASES)))
def test_oks_functions_match_reference_golden(golden, ci):
    g = golden("nms_ref.npz")
    people, seed, thr, vthr = ggn.NMS_CASES[ci]
    kpts, areas, scores = ggn.nms_people(seed, people)
    flat = kpts.reshape(people, -1)
    iou = nms.oks_iou(flat[0], flat, areas[0], areas, None, vthr)
    assert iou.dtype == np.float32
    assert np.array_equal(iou, g[f"nms{ci}_iou"])          # bit-exact, incl. the pairwise sum
    assert np.array_equal(nms.oks_nms(flat, areas, scores, thr, vis_thr=vthr), g[f"nms{ci}_keep"])
    assert np.array_equal(nms.soft_oks_nms(flat, areas, scores, thr, 20, vis_thr=vthr),
                          g[f"nms{ci}_soft"])


@pytest.mark.parametrize("ci", range(len(ggn.EVAL_CASES)))
def test_evaluator_rescoring_and_nms_match_reference_golden(golden, ci):
    g = golden("nms_ref.npz")
    images, max_people, seed, soft = ggn.EVAL_CASES[ci]
    records = ggn.eval_records(images, max_people, seed)
    kept = nms.evaluate_records(records, vis_thr=0.2, oks_thr=0.9, use_nms=True, soft_nms=soft,
                                sigmas=nms.COCO_SIGMAS)
    assert np.array_equal([len(k) for k in kept], g[f"eval{ci}_counts"])
    assert np.array_equal([b for k in kept for b, _ in k], g[f"eval{ci}_bbox_ids"])
    assert np.array_equal(np.asarray([s for k in kept for _, s in k], np.float32),
                          g[f"eval{ci}_scores"])       # float32 rescoring: bit-exact


@pytest.mark.parametrize("ci", range(len(ggn.NMS_CASES)))
def test_oks_functions_match_reference_golden_for_list_records(golden, ci):
    """Records holding Python floats (the inferencer's `.tolist()` format): float64."""
    g = golden("nms_ref.npz")
    people, seed, thr, vthr = ggn.NMS_CASES[ci]
    kpts, areas, scores = ggn.nms_people(seed, people)
    flat = kpts.reshape(people, -1).astype(np.float64)
    a64, s64 = areas.astype(np.float64), scores.astype(np.float64)
    iou = nms.oks_iou(flat[0], flat, a64[0], a64, None, vthr, dtype=np.float64)
    assert iou.dtype == np.float32
    assert np.array_equal(iou, g[f"nmsL{ci}_iou"])
    assert np.array_equal(nms.oks_nms(flat, a64, s64, thr, vis_thr=vthr, dtype=np.float64),
                          g[f"nmsL{ci}_keep"])
    assert np.array_equal(nms.soft_oks_nms(flat, a64, s64, thr, 20, vis_thr=vthr,
                                           dtype=np.float64), g[f"nmsL{ci}_soft"])


@pytest.mark.parametrize("ci", range(len(ggn.EVAL_CASES)))
def test_evaluator_matches_reference_golden_for_list_records(golden, ci):
    g = golden("nms_ref.npz")
    images, max_people, seed, soft = ggn.EVAL_CASES[ci]
    records = ggn.eval_records(images, max_people, seed, as_lists=True)
    assert nms.records_dtype(records) == np.float64
    kept = nms.evaluate_records(records, vis_thr=0.2, oks_thr=0.9, use_nms=True, soft_nms=soft,
                                sigmas=nms.COCO_SIGMAS)
    assert np.array_equal([len(k) for k in kept], g[f"evalL{ci}_counts"])
    assert np.array_equal([b for k in kept for b, _ in k], g[f"evalL{ci}_bbox_ids"])
    got = np.asarray([s for k in kept for _, s in k])
    assert got.dtype == np.float64
    assert np.array_equal(got, g[f"evalL{ci}_scores"])     # float64 rescoring: bit-exact


def test_pairwise_sum_is_numpy_sum():
    rng = np.random.RandomState(0)
    for n in list(range(0, 40)) + [64, 100, 127]:
        a = np.exp(-rng.uniform(0, 5, n))
        assert nms._pairwise_sum(a) == (np.sum(a) if n else 0.0)


def test_sort_and_unique_keeps_first_of_each_id():
    ids = np.array([5, 3, 5, 1, 3, 9])
    assert np.array_equal(nms.sort_and_unique(ids), [3, 1, 0, 5])


@pytest.mark.needs_reference
def test_live_reference_agrees_on_fresh_seeds():
    ref = ref_loader.load()
    for seed in range(100, 110):
        people = 5 + 7 * (seed % 5)
        kpts, areas, scores = ggn.nms_people(seed, people)
        db = ggn._kpts_db(kpts, areas, scores)
        flat = kpts.reshape(people, -1)
        assert np.array_equal(nms.oks_nms(flat, areas, scores, 0.8), ref.nms.oks_nms(db, 0.8))
        assert np.array_equal(nms.soft_oks_nms(flat, areas, scores, 0.8, 20),
                              ref.nms.soft_oks_nms(db, 0.8, max_dets=20))


@pytest.mark.needs_reference
def test_live_reference_agrees_for_other_joint_counts_and_vis_thr():
    """K = 5 (numpy sums fewer than 8 terms in order), K = 21, custom sigmas, and the
    detection-only vis_thr selection (nms.py:64)."""
    ref = ref_loader.load()
    rng = np.random.RandomState(3)
    for k in (5, 8, 21):
        sig = rng.uniform(0.02, 0.11, k)
        for seed in range(3):
            people = 9 + 4 * seed
            kpts, areas, scores = ggn.nms_people(200 + seed, people, k)
            flat = kpts.reshape(people, -1)
            db = ggn._kpts_db(kpts, areas, scores)
            for vthr in (None, 0.4):
                want = ref.nms.oks_iou(flat[0], flat, areas[0], areas, sig, vthr)
                assert np.array_equal(nms.oks_iou(flat[0], flat, areas[0], areas, sig, vthr), want)
                assert np.array_equal(nms.oks_nms(flat, areas, scores, 0.7, sig, vthr),
                                      ref.nms.oks_nms(db, 0.7, sigmas=sig, vis_thr=vthr))
                assert np.array_equal(nms.soft_oks_nms(flat, areas, scores, 0.7, 6, sig, vthr),
                                      ref.nms.soft_oks_nms(db, 0.7, max_dets=6, sigmas=sig,
                                                           vis_thr=vthr))
