"""Oracle (test infrastructure): OKS rescoring + OKS NMS / soft OKS NMS.

numpy restatement of

* the rescoring loop of ``TopDownEvaluator.eval``
  (mindpose/engine/evaluator/topdown_evaluator.py:93-110),
* ``_sort_and_unique_bboxes`` (topdown_evaluator.py:139-148),
* ``oks_iou`` / ``oks_nms`` / ``_rescore`` / ``soft_oks_nms`` (mindpose/utils/nms.py:7-190).

PINNED: ``oracle/gen_golden_nms.py`` runs the unmodified reference functions (imported
from /root/reference) on seeded inputs; ``tests/test_oracle_nms.py`` checks this module
against the stored outputs.

Arithmetic notes (NumPy >= 2 promotion rules, the rules of the numpy this image and the
golden vectors use; under NumPy 1.x ``0 + np.float32`` is float64 and the rescored
scores differ in the last float32 bit):

* rescoring: float32 running sum over the joints whose score is > float32(vis_thr), in
  joint order, float32 divide by the count, float32 multiply by the box score;
* ``oks_iou``: ``dx**2 + dy**2`` in float32; ``/ key_vars`` promotes to float64; the
  area term ``(a_g + a_d) / 2`` is float32 and becomes float64 when ``np.spacing(1)``
  (a float64 scalar) is added; ``np.sum(np.exp(-e))`` is numpy's pairwise sum of a
  contiguous float64 vector (eight running sums for n >= 8, the remainder added in
  order); the result is stored into a float32 array;
* ``vis_thr`` quirk (nms.py:64): ``list(vg > t) and list(vd > t)`` is the SECOND list
  whenever the first is non-empty, i.e. only the detection's visibilities select joints;
* ``oks_ovr <= thr`` compares float32 values with float32(thr);
* ``scores.argsort()[::-1]`` uses an unstable sort: the order of EQUAL scores is not
  defined by the reference.  Canonical rule here and in the CUDA kernel: a stable sort
  reversed, i.e. (score descending, current position descending);
* soft NMS rescoring is float32 throughout, ``np.exp`` on float32 (numpy's SIMD exp,
  <= 2.5 ulp): the kernel rounds a float64 exp to float32 instead, so rescored values may
  differ by an ulp; only the keep order leaves the function.
"""
import numpy as np

F32 = np.float32
COCO_SIGMAS = np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07,
                        .87, .87, .89, .89]) / 10.0


def rescore(preds, box_scores, vis_thr):
    """preds f32 [P,K,3], box_scores f32 [P] -> rescored f32 [P]
    (topdown_evaluator.py:96-110)."""
    preds = np.asarray(preds, dtype=F32)
    out = np.zeros(len(preds), dtype=F32)
    thr = F32(vis_thr)
    for p in range(len(preds)):
        acc = F32(0)
        cnt = 0
        for j in range(preds.shape[1]):
            t = preds[p, j, 2]
            if t > thr:
                acc = F32(acc + t)
                cnt += 1
        if cnt:
            acc = F32(acc / F32(cnt))
        out[p] = F32(acc * F32(box_scores[p]))
    return out


def sort_and_unique(bbox_ids):
    """Positions kept by ``_sort_and_unique_bboxes`` (topdown_evaluator.py:139-148):
    stable sort by bbox_id, then of each run of equal ids the FIRST survives."""
    order = np.argsort(np.asarray(bbox_ids), kind="stable")
    keep = [order[0]] if len(order) else []
    for a, b in zip(order[:-1], order[1:]):
        if bbox_ids[b] != bbox_ids[a]:
            keep.append(b)
    return np.asarray(keep, dtype=np.int64)


def _pairwise_sum(a):
    """numpy's pairwise_sum for a contiguous vector shorter than the 128-element block."""
    n = len(a)
    if n < 8:
        res = 0.0
        for x in a:
            res += x
        return res
    r = [a[i] for i in range(8)]
    i = 8
    while i < n - (n % 8):
        for j in range(8):
            r[j] += a[i + j]
        i += 8
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    while i < n:
        res += a[i]
        i += 1
    return res


def oks_iou(g, d, a_g, a_d, sigmas=None, vis_thr=None):
    """g f32 [3K], d f32 [n,3K], a_g f32, a_d f32 [n] -> f32 [n] (nms.py:7-69)."""
    sig = np.asarray(COCO_SIGMAS if sigmas is None else sigmas, dtype=np.float64)
    key_vars = (sig * 2) ** 2
    g = np.asarray(g, dtype=F32)
    d = np.asarray(d, dtype=F32).reshape(-1, g.size)
    xg, yg = g[0::3], g[1::3]
    out = np.zeros(len(d), dtype=F32)
    for n in range(len(d)):
        dx = (d[n, 0::3] - xg).astype(F32)
        dy = (d[n, 1::3] - yg).astype(F32)
        sq = (dx * dx + dy * dy).astype(F32)
        area = np.float64(F32(F32(F32(a_g) + F32(a_d[n])) / F32(2))) + np.spacing(1)
        e = sq.astype(np.float64) / key_vars / area / 2
        if vis_thr is not None:
            e = e[d[n, 2::3] > F32(vis_thr)]
        out[n] = _pairwise_sum(np.exp(-e)) / len(e) if e.size else 0.0
    return out


def _argsort_desc(scores):
    """Canonical ``scores.argsort()[::-1]``: stable ascending sort, reversed."""
    return np.argsort(scores, kind="stable")[::-1]


def oks_nms(kpts, areas, scores, thr, sigmas=None, vis_thr=None):
    """kpts f32 [P,3K], areas f32 [P], scores f32 [P] -> kept indices (nms.py:72-111)."""
    if len(scores) == 0:
        return np.zeros(0, dtype=np.int64)
    order = _argsort_desc(np.asarray(scores))
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(i)
        ovr = oks_iou(kpts[i], kpts[order[1:]], areas[i], areas[order[1:]], sigmas, vis_thr)
        order = order[np.where(ovr <= F32(thr))[0] + 1]
    return np.asarray(keep, dtype=np.int64)


def soft_oks_nms(kpts, areas, scores, thr, max_dets=20, sigmas=None, vis_thr=None):
    """-> kept indices, at most max_dets (nms.py:141-190; gaussian rescoring :114-138)."""
    if len(scores) == 0:
        return np.zeros(0, dtype=np.int64)
    scores = np.asarray(scores, dtype=F32)
    order = _argsort_desc(scores)
    scores = scores[order]
    keep = []
    while order.size > 0 and len(keep) < max_dets:
        i = order[0]
        ovr = oks_iou(kpts[i], kpts[order[1:]], areas[i], areas[order[1:]], sigmas, vis_thr)
        order = order[1:]
        w = np.exp((-(ovr * ovr).astype(F32) / F32(thr)).astype(np.float64)).astype(F32)
        scores = (scores[1:] * w).astype(F32)
        tmp = _argsort_desc(scores)
        order = order[tmp]
        scores = scores[tmp]
        keep.append(i)
    return np.asarray(keep, dtype=np.int64)


def evaluate_records(records, vis_thr, oks_thr, use_nms=True, soft_nms=False, sigmas=None):
    """``TopDownEvaluator.eval`` up to the result file (topdown_evaluator.py:78-121):
    group the inference records by image (first-seen order), sort / de-duplicate by
    bbox_id, rescore, NMS.  -> per image: list of (bbox_id, rescored score) in keep order."""
    by_image = {}
    for rec in records:
        by_image.setdefault(rec["image_path"].split("/")[-1], []).append(rec)
    out = []
    for recs in by_image.values():
        ids = np.asarray([r["bbox_id"] for r in recs])
        recs = [recs[i] for i in sort_and_unique(ids)]
        preds = np.stack([np.asarray(r["pred"], dtype=F32) for r in recs])
        boxes = np.stack([np.asarray(r["box"], dtype=F32) for r in recs])
        scores = rescore(preds, boxes[:, 5], vis_thr)
        if use_nms:
            fn = soft_oks_nms if soft_nms else oks_nms
            keep = fn(preds.reshape(len(recs), -1), boxes[:, 4], scores, oks_thr, sigmas=sigmas)
        else:
            keep = np.arange(len(recs))
        out.append([(int(recs[i]["bbox_id"]), scores[i]) for i in keep])
    return out
