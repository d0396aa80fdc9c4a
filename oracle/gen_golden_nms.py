"""Golden vectors for row N4 (OKS rescoring + OKS NMS), made by the UNMODIFIED reference.

    python -m oracle.gen_golden_nms          # writes tests/golden/nms_ref.npz

Runs in the build container only (needs /root/reference).  ``mindpose/utils/nms.py`` is
pure numpy and is imported as it is.  The rescoring loop lives inside
``TopDownEvaluator.eval`` (mindpose/engine/evaluator/topdown_evaluator.py:68-123); that
module is imported unchanged too -- ``pycocotools`` (not installed here) is replaced by an
empty stub because only ``eval``'s list handling is exercised -- and ``eval`` runs on a
bare instance whose ``_write_coco_keypoint_results`` captures the kept people and stops.
"""
import importlib
import os
import sys
import types

import numpy as np

from oracle import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (people, seed, oks_thr, iou_vis_thr)
NMS_CASES = [
    (1, 0, 0.9, None), (2, 1, 0.9, None), (3, 2, 0.5, None), (8, 3, 0.9, None),
    (17, 4, 0.9, None), (17, 5, 0.6, 0.3), (40, 6, 0.9, None), (40, 7, 0.7, 0.5),
    (100, 8, 0.9, None), (100, 9, 0.8, None), (200, 10, 0.9, None), (33, 11, 0.3, None),
]
# (images, max people per image, seed, use soft nms)
EVAL_CASES = [(6, 12, 20, False), (5, 30, 21, False), (4, 25, 22, True), (3, 60, 23, True)]


def nms_people(seed, people, k=17):
    """kpts f32 [P,k,3], areas f32 [P], box scores f32 [P]: a few base poses, each with
    several jittered near-duplicates, so that OKS covers 0.2 .. 0.99."""
    rng = np.random.RandomState(seed)
    base_n = max(1, people // 4)
    centers = rng.uniform(50, 400, size=(base_n, 2))
    sizes = rng.uniform(40, 160, size=base_n)
    shapes = rng.uniform(-0.5, 0.5, size=(base_n, k, 2))
    kpts = np.zeros((people, k, 3), np.float32)
    areas = np.zeros(people, np.float32)
    for p in range(people):
        b = rng.randint(base_n)
        jitter = rng.choice([0.01, 0.03, 0.08, 0.2]) * sizes[b]
        kpts[p, :, :2] = centers[b] + shapes[b] * sizes[b] + rng.normal(0, jitter, size=(k, 2))
        kpts[p, :, 2] = rng.uniform(0, 1, size=k)
        areas[p] = sizes[b] ** 2 * rng.uniform(0.8, 1.2)
    scores = rng.uniform(0.05, 1.0, size=people).astype(np.float32)
    return kpts, areas, scores


def _kpts_db(kpts, areas, scores, as_lists=False):
    """as_lists: Python lists / floats, what records look like after the inferencer's
    ``.tolist()`` (topdown_inferencer.py:135-140) -- numpy then computes in float64."""
    if as_lists:
        return [dict(keypoints=kpts[p].tolist(), area=float(areas[p]), score=float(scores[p]))
                for p in range(len(kpts))]
    return [dict(keypoints=kpts[p], area=areas[p], score=scores[p]) for p in range(len(kpts))]


def _load_evaluator():
    for name in ("pycocotools", "pycocotools.coco", "pycocotools.cocoeval"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["pycocotools.coco"].COCO = object
    sys.modules["pycocotools.cocoeval"].COCOeval = object
    ref_loader.load()
    ref_loader._skeleton("mindpose.engine", "mindpose/engine")
    ref_loader._skeleton("mindpose.engine.evaluator", "mindpose/engine/evaluator")
    return importlib.import_module("mindpose.engine.evaluator.topdown_evaluator")


class _Captured(Exception):
    pass


def run_reference_eval(records, name2id, cfg, num_joints=17):
    """``TopDownEvaluator.eval`` up to the result file: -> list (per image, in first-seen
    order) of lists of (bbox_id, rescored score) of the people that are written out."""
    mod = _load_evaluator()
    ev = object.__new__(mod.TopDownEvaluator)
    ev.name2id = name2id
    ev.num_joints = num_joints
    ev._evaluation_cfg = cfg
    got = {}

    def capture(valid_kpts, path):
        got["kept"] = [[(int(p["bbox_id"]), p["score"]) for p in img] for img in valid_kpts]
        raise _Captured()

    ev._write_coco_keypoint_results = capture
    ev.result_path = "/dev/null"
    try:
        ev.eval(records)
    except _Captured:
        pass
    return got["kept"]


def eval_records(images, max_people, seed, k=17, as_lists=False):
    """Inference records: pred [k,3], box [6], bbox_id, image_path.  as_lists=True gives them
    exactly as the top-down inferencer emits them (``pred.tolist()``, ``box.tolist()``,
    topdown_inferencer.py:135-140); False keeps float32 ndarrays.
    Some bbox_ids are duplicated (flip-test style repeats)."""
    rng = np.random.RandomState(seed)
    records = []
    bbox_id = 0
    for im in range(images):
        people = rng.randint(1, max_people + 1)
        kpts, areas, scores = nms_people(seed * 100 + im, people, k)
        for p in range(people):
            box = np.zeros(6, np.float32)
            box[0:2] = kpts[p, :, :2].mean(0)
            box[2:4] = np.sqrt(areas[p]) / 200.0
            box[4] = areas[p]
            box[5] = scores[p]
            rec = dict(pred=kpts[p], box=box, image_path=f"/data/img_{im:04d}.jpg", bbox_id=bbox_id)
            records.append(rec)
            if rng.random_sample() < 0.15:          # a repeated box: must be dropped
                dup = dict(rec)
                dup["pred"] = (kpts[p] + np.float32(1.0)).astype(np.float32)
                records.append(dup)
            bbox_id += 1
    order = rng.permutation(len(records))
    records = [records[i] for i in order]
    if as_lists:
        records = [dict(r, pred=r["pred"].tolist(), box=r["box"].tolist()) for r in records]
    return records


def main():
    ref = ref_loader.load()
    out = {}
    for ci, (people, seed, thr, vthr) in enumerate(NMS_CASES):
        kpts, areas, scores = nms_people(seed, people)
        db = _kpts_db(kpts, areas, scores)
        sig = None  # the function's own COCO default
        out[f"nms{ci}_keep"] = np.asarray(ref.nms.oks_nms(db, thr, sigmas=sig, vis_thr=vthr),
                                          dtype=np.int64)
        out[f"nms{ci}_soft"] = np.asarray(
            ref.nms.soft_oks_nms(db, thr, max_dets=20, sigmas=sig, vis_thr=vthr), dtype=np.int64)
        flat = kpts.reshape(people, -1)
        out[f"nms{ci}_iou"] = ref.nms.oks_iou(flat[0], flat, areas[0], areas, sig, vthr)
        dbl = _kpts_db(kpts, areas, scores, as_lists=True)      # float64 arithmetic
        out[f"nmsL{ci}_keep"] = np.asarray(ref.nms.oks_nms(dbl, thr, sigmas=sig, vis_thr=vthr),
                                           dtype=np.int64)
        out[f"nmsL{ci}_soft"] = np.asarray(
            ref.nms.soft_oks_nms(dbl, thr, max_dets=20, sigmas=sig, vis_thr=vthr), dtype=np.int64)
        f64 = flat.astype(np.float64)
        out[f"nmsL{ci}_iou"] = ref.nms.oks_iou(f64[0], f64, float(areas[0]),
                                               areas.astype(np.float64), sig, vthr)
    sigmas = (np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07,
                        .87, .87, .89, .89]) / 10.0).tolist()
    for ci, (images, max_people, seed, soft) in enumerate(EVAL_CASES):
        records = eval_records(images, max_people, seed)
        name2id = {f"img_{im:04d}.jpg": 1000 + im for im in range(images)}
        cfg = dict(vis_thr=0.2, oks_thr=0.9, use_nms=True, soft_nms=soft, sigmas=sigmas)
        kept = run_reference_eval(records, name2id, cfg)
        out[f"eval{ci}_counts"] = np.asarray([len(img) for img in kept], dtype=np.int64)
        out[f"eval{ci}_bbox_ids"] = np.asarray([b for img in kept for b, _ in img], dtype=np.int64)
        out[f"eval{ci}_scores"] = np.asarray([s for img in kept for _, s in img], dtype=np.float32)
        # the same records as lists (the inferencer's own format): float64 rescoring and NMS
        kept = run_reference_eval(eval_records(images, max_people, seed, as_lists=True), name2id,
                                  cfg)
        out[f"evalL{ci}_counts"] = np.asarray([len(img) for img in kept], dtype=np.int64)
        out[f"evalL{ci}_bbox_ids"] = np.asarray([b for img in kept for b, _ in img], dtype=np.int64)
        out[f"evalL{ci}_scores"] = np.asarray([s for img in kept for _, s in img], dtype=np.float64)
    path = os.path.join(ROOT, "tests", "golden", "nms_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in list(out.items())[:6]})


if __name__ == "__main__":
    main()
