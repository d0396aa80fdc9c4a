"""Row N4 on the device (pc_oks_nms) against the reference's golden vectors and the oracle
(B200 only).  Bars: kept indices and float32 rescored scores bit-exact."""
import numpy as np
import pytest
import torch

from mindpose_b200 import nms as dnms
from oracle import gen_golden_nms as ggn
from oracle import nms as onms

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ci", range(len(ggn.NMS_CASES)))
def test_oks_nms_matches_reference_golden(cuda_device, golden, ci):
    g = golden("nms_ref.npz")
    people, seed, thr, vthr = ggn.NMS_CASES[ci]
    kpts, areas, scores = ggn.nms_people(seed, people)
    db = ggn._kpts_db(kpts, areas, scores)
    assert np.array_equal(dnms.oks_nms(db, thr, vis_thr=vthr, device=cuda_device), g[f"nms{ci}_keep"])
    assert np.array_equal(dnms.soft_oks_nms(db, thr, max_dets=20, vis_thr=vthr, device=cuda_device),
                          g[f"nms{ci}_soft"])


@pytest.mark.parametrize("ci", range(len(ggn.EVAL_CASES)))
def test_evaluator_flow_matches_reference_golden(cuda_device, golden, ci):
    g = golden("nms_ref.npz")
    images, max_people, seed, soft = ggn.EVAL_CASES[ci]
    records = ggn.eval_records(images, max_people, seed)
    cfg = dict(vis_thr=0.2, oks_thr=0.9, use_nms=True, soft_nms=soft, sigmas=onms.COCO_SIGMAS)
    kept = dnms.evaluate_records(records, cfg, device=cuda_device)
    assert np.array_equal([len(k) for k in kept], g[f"eval{ci}_counts"])
    assert np.array_equal([r["bbox_id"] for k in kept for r in k], g[f"eval{ci}_bbox_ids"])
    assert np.array_equal(np.asarray([r["score"] for k in kept for r in k], np.float32),
                          g[f"eval{ci}_scores"])


@pytest.mark.parametrize("ci", range(len(ggn.NMS_CASES)))
def test_oks_nms_list_records_match_reference_golden(cuda_device, golden, ci):
    """Python-float records (the inferencer's `.tolist()` format): numpy, and so
    pc_oks_nms_f64, computes in float64."""
    g = golden("nms_ref.npz")
    people, seed, thr, vthr = ggn.NMS_CASES[ci]
    kpts, areas, scores = ggn.nms_people(seed, people)
    db = ggn._kpts_db(kpts, areas, scores, as_lists=True)
    assert np.array_equal(dnms.oks_nms(db, thr, vis_thr=vthr, device=cuda_device), g[f"nmsL{ci}_keep"])
    assert np.array_equal(dnms.soft_oks_nms(db, thr, max_dets=20, vis_thr=vthr, device=cuda_device),
                          g[f"nmsL{ci}_soft"])


@pytest.mark.parametrize("ci", range(len(ggn.EVAL_CASES)))
def test_evaluator_flow_list_records_match_reference_golden(cuda_device, golden, ci):
    """The records exactly as TopDownHeatMapInferencer emits them: rescored float64 scores
    bit-exact vs TopDownEvaluator.eval of the unmodified reference."""
    g = golden("nms_ref.npz")
    images, max_people, seed, soft = ggn.EVAL_CASES[ci]
    records = ggn.eval_records(images, max_people, seed, as_lists=True)
    cfg = dict(vis_thr=0.2, oks_thr=0.9, use_nms=True, soft_nms=soft, sigmas=onms.COCO_SIGMAS)
    kept = dnms.evaluate_records(records, cfg, device=cuda_device)
    assert np.array_equal([len(k) for k in kept], g[f"evalL{ci}_counts"])
    assert np.array_equal([r["bbox_id"] for k in kept for r in k], g[f"evalL{ci}_bbox_ids"])
    got = np.asarray([r["score"] for k in kept for r in k])
    assert got.dtype == np.float64
    assert np.array_equal(got, g[f"evalL{ci}_scores"])


def test_batched_nms_matches_oracle_with_ties_and_empty_images(cuda_device):
    """Many images in one launch: empty images, one person, tied scores (canonical rule),
    up to 300 people; hard and soft."""
    rng = np.random.RandomState(3)
    counts = [0, 1, 7, 0, 64, 300, 2, 129]
    ks, ars, scs = [], [], []
    for i, c in enumerate(counts):
        k, a, s = ggn.nms_people(40 + i, max(c, 1))
        if i in (2, 4):
            s = np.round(s * 4).astype(np.float32) / 4          # heavy ties
        ks.append(k[:c]), ars.append(a[:c]), scs.append(s[:c])
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    kp = np.concatenate(ks).astype(np.float32)
    ar = np.concatenate(ars).astype(np.float32)
    sc = np.concatenate(scs).astype(np.float32)
    for soft in (False, True):
        score = torch.from_numpy(sc.copy()).to(cuda_device)
        keep, num = dnms.rescore_and_nms(
            torch.from_numpy(kp).to(cuda_device), torch.from_numpy(ar).to(cuda_device), score,
            torch.from_numpy(off).to(cuda_device), max(counts), oks_thr=0.8,
            rescore_vis_thr=0.2, soft=soft, max_dets=20)
        keep, num = keep.cpu().numpy(), num.cpu().numpy()
        for i, c in enumerate(counts):
            lo = off[i]
            want_s = onms.rescore(ks[i], scs[i], 0.2) if c else np.zeros(0, np.float32)
            assert np.array_equal(score.cpu().numpy()[lo:lo + c], want_s)
            fn = onms.soft_oks_nms if soft else onms.oks_nms
            args = (ks[i].reshape(c, 51), ars[i], want_s, 0.8) + ((20,) if soft else ())
            want = fn(*args) if c else np.zeros(0, np.int64)
            assert num[i] == len(want)
            assert np.array_equal(keep[lo:lo + num[i]], want)
            assert np.all(keep[lo + num[i]:lo + c] == -1)


def test_use_nms_false_and_too_many_people(cuda_device):
    k, a, s = ggn.nms_people(1, 9)
    off = torch.tensor([0, 9], dtype=torch.int32, device=cuda_device)
    t = lambda x: torch.from_numpy(x).to(cuda_device)  # noqa: E731
    keep, num = dnms.rescore_and_nms(t(k), t(a), t(s.copy()), off, 9, oks_thr=0.9, use_nms=False)
    assert num.item() == 9 and np.array_equal(keep.cpu().numpy(), np.arange(9))
    keep, num = dnms.rescore_and_nms(t(k), t(a), t(s.copy()), off, 4, oks_thr=0.9)
    assert num.item() == -1                      # host under-declared max_people_per_image
    with pytest.raises(ValueError):
        dnms.rescore_and_nms(t(k), t(a), t(s.copy()), off, 5000, oks_thr=0.9)
    with pytest.raises(ValueError):
        dnms.rescore_and_nms(t(k).cpu(), t(a), t(s.copy()), off, 9, oks_thr=0.9)
    assert dnms.oks_nms([], 0.9) == []


@pytest.mark.parametrize("k", [5, 8, 21])
@pytest.mark.parametrize("vthr", [None, 0.4])
def test_other_joint_counts_custom_sigmas_and_vis_thr(cuda_device, k, vthr):
    """K != 17: fewer than 8 exp terms are summed in order (numpy's pairwise rule), custom
    sigmas, and the detection-only visibility selection of oks_iou."""
    rng = np.random.RandomState(k)
    sig = rng.uniform(0.02, 0.11, k)
    for seed in range(3):
        people = 9 + 4 * seed
        kpts, areas, scores = ggn.nms_people(200 + seed, people, k)
        db = ggn._kpts_db(kpts, areas, scores)
        flat = kpts.reshape(people, -1)
        assert np.array_equal(dnms.oks_nms(db, 0.7, sigmas=sig, vis_thr=vthr, device=cuda_device),
                              onms.oks_nms(flat, areas, scores, 0.7, sig, vthr))
        assert np.array_equal(
            dnms.soft_oks_nms(db, 0.7, max_dets=6, sigmas=sig, vis_thr=vthr, device=cuda_device),
            onms.soft_oks_nms(flat, areas, scores, 0.7, 6, sig, vthr))
