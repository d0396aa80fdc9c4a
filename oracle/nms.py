"""Oracle (test infrastructure): OKS rescoring + OKS NMS / soft OKS NMS.

numpy restatement of

* the rescoring loop of ``TopDownEvaluator.eval``
  (mindpose/engine/evaluator/topdown_evaluator.py:93-110),
* ``_sort_and_unique_bboxes`` (topdown_evaluator.py:139-148),
* ``oks_iou`` / ``oks_nms`` / ``_rescore`` / ``soft_oks_nms`` (mindpose/utils/nms.py:7-190).

PINNED: ``oracle/gen_golden_nms.py`` runs the unmodified reference functions (imported
from /root/reference) on seeded inputs; ``tests/test_oracle_nms.py`` checks this module
against the stored outputs.

Precision.  The reference computes in whatever dtype its records hold.  The inferencer emits
``pred.tolist()`` / ``box.tolist()`` (topdown_inferencer.py:135-140): Python floats, so the
evaluator's rescoring, ``dx**2 + dy**2``, the areas and the sort keys run in float64
(``dtype=np.float64`` below, the CUDA entry ``pc_oks_nms_f64``).  Records holding float32
ndarrays stay in float32 where numpy does (``dtype=np.float32``, ``pc_oks_nms``).  The notes
below are written for float32; for float64 read "float64" for every "float32" except the OKS
values themselves (``ious = np.zeros(..., dtype=np.float32)``, nms.py:56), the comparisons with
``thr`` and the soft-NMS weight ``np.exp(-(overlap**2) / thr)``, which stay float32.

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


def rescore(preds, box_scores, vis_thr, dtype=F32):
    """preds [P,K,3], box_scores [P] -> rescored [P], all `dtype`
    (topdown_evaluator.py:96-110)."""
    T = np.dtype(dtype).type
    preds = np.asarray(preds, dtype=T)
    out = np.zeros(len(preds), dtype=T)
    thr = T(vis_thr)
    for p in range(len(preds)):
        acc = T(0)
        cnt = 0
        for j in range(preds.shape[1]):
            t = preds[p, j, 2]
            if t > thr:
                acc = T(acc + t)
                cnt += 1
        if cnt:
            acc = T(acc / T(cnt))
        out[p] = T(acc * T(box_scores[p]))
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


def oks_iou(g, d, a_g, a_d, sigmas=None, vis_thr=None, dtype=F32):
    """g [3K], d [n,3K], a_g, a_d [n] (all `dtype`) -> f32 [n] (nms.py:7-69)."""
    T = np.dtype(dtype).type
    sig = np.asarray(COCO_SIGMAS if sigmas is None else sigmas, dtype=np.float64)
    key_vars = (sig * 2) ** 2
    g = np.asarray(g, dtype=T)
    d = np.asarray(d, dtype=T).reshape(-1, g.size)
    xg, yg = g[0::3], g[1::3]
    out = np.zeros(len(d), dtype=F32)
    for n in range(len(d)):
        dx = (d[n, 0::3] - xg).astype(T)
        dy = (d[n, 1::3] - yg).astype(T)
        sq = (dx * dx + dy * dy).astype(T)
        area = np.float64(T(T(T(a_g) + T(a_d[n])) / T(2))) + np.spacing(1)
        e = sq.astype(np.float64) / key_vars / area / 2
        if vis_thr is not None:
            e = e[d[n, 2::3] > T(vis_thr)]
        out[n] = _pairwise_sum(np.exp(-e)) / len(e) if e.size else 0.0
    return out


def _argsort_desc(scores):
    """Canonical ``scores.argsort()[::-1]``: stable ascending sort, reversed."""
    return np.argsort(scores, kind="stable")[::-1]


def oks_nms(kpts, areas, scores, thr, sigmas=None, vis_thr=None, dtype=F32):
    """kpts [P,3K], areas [P], scores [P] (all `dtype`) -> kept indices (nms.py:72-111)."""
    if len(scores) == 0:
        return np.zeros(0, dtype=np.int64)
    kpts, areas = np.asarray(kpts, dtype=dtype), np.asarray(areas, dtype=dtype)
    order = _argsort_desc(np.asarray(scores, dtype=dtype))
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(i)
        ovr = oks_iou(kpts[i], kpts[order[1:]], areas[i], areas[order[1:]], sigmas, vis_thr,
                      dtype)
        order = order[np.where(ovr <= F32(thr))[0] + 1]
    return np.asarray(keep, dtype=np.int64)


def soft_oks_nms(kpts, areas, scores, thr, max_dets=20, sigmas=None, vis_thr=None, dtype=F32):
    """-> kept indices, at most max_dets (nms.py:141-190; gaussian rescoring :114-138)."""
    if len(scores) == 0:
        return np.zeros(0, dtype=np.int64)
    T = np.dtype(dtype).type
    kpts, areas = np.asarray(kpts, dtype=T), np.asarray(areas, dtype=T)
    scores = np.asarray(scores, dtype=T)
    order = _argsort_desc(scores)
    scores = scores[order]
    keep = []
    while order.size > 0 and len(keep) < max_dets:
        i = order[0]
        ovr = oks_iou(kpts[i], kpts[order[1:]], areas[i], areas[order[1:]], sigmas, vis_thr, T)
        order = order[1:]
        # the weight is float32 (the OKS values are); the product takes the scores' precision
        w = np.exp((-(ovr * ovr).astype(F32) / F32(thr)).astype(np.float64)).astype(F32)
        scores = (scores[1:] * w.astype(T)).astype(T)
        tmp = _argsort_desc(scores)
        order = order[tmp]
        scores = scores[tmp]
        keep.append(i)
    return np.asarray(keep, dtype=np.int64)


def records_dtype(records):
    """float32 when every record holds float32 ndarrays, else float64 (Python floats from
    ``.tolist()``, float64 arrays): the dtype numpy computes in for such records."""
    ok = all(np.asarray(r["pred"]).dtype == F32 and np.asarray(r["box"]).dtype == F32
             for r in records)
    return F32 if ok else np.float64


def evaluate_records(records, vis_thr, oks_thr, use_nms=True, soft_nms=False, sigmas=None):
    """``TopDownEvaluator.eval`` up to the result file (topdown_evaluator.py:78-121):
    group the inference records by image (first-seen order), sort / de-duplicate by
    bbox_id, rescore, NMS.  -> per image: list of (bbox_id, rescored score) in keep order."""
    T = records_dtype(records)
    by_image = {}
    for rec in records:
        by_image.setdefault(rec["image_path"].split("/")[-1], []).append(rec)
    out = []
    for recs in by_image.values():
        ids = np.asarray([r["bbox_id"] for r in recs])
        recs = [recs[i] for i in sort_and_unique(ids)]
        preds = np.stack([np.asarray(r["pred"], dtype=T) for r in recs])
        boxes = np.stack([np.asarray(r["box"], dtype=T) for r in recs])
        scores = rescore(preds, boxes[:, 5], vis_thr, T)
        if use_nms:
            fn = soft_oks_nms if soft_nms else oks_nms
            keep = fn(preds.reshape(len(recs), -1), boxes[:, 4], scores, oks_thr, sigmas=sigmas,
                      dtype=T)
        else:
            keep = np.arange(len(recs))
        out.append([(int(recs[i]["bbox_id"]), scores[i]) for i in keep])
    return out
