"""OKS rescoring + OKS NMS on the device.

Host mirror of ``mindpose/utils/nms.py`` (``oks_nms`` / ``soft_oks_nms`` keep their
signatures: a list of ``{"keypoints", "area", "score"}`` dicts in, kept indices out) and
of the rescoring / NMS block of ``TopDownEvaluator.eval``
(mindpose/engine/evaluator/topdown_evaluator.py:78-121) for a whole evaluation at once.
Everything numeric runs in ``pc_oks_nms``; the host only groups records by image and
sorts / de-duplicates them by ``bbox_id`` (list handling, as in the reference).
"""
import ctypes
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

COCO_SIGMAS = (np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07,
                         .87, .87, .89, .89]) / 10.0)


def rescore_and_nms(kpts: torch.Tensor, area: torch.Tensor, score: torch.Tensor,
                    image_offset: torch.Tensor, max_people_per_image: int, *,
                    oks_thr: float, rescore_vis_thr: Optional[float] = None, use_nms: bool = True,
                    soft: bool = False, max_dets: int = 20, sigmas=None,
                    iou_vis_thr: Optional[float] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """One ``pc_oks_nms`` / ``pc_oks_nms_f64`` launch over all images.

    kpts [P,K,3], area [P], score [P] (rescored IN PLACE when ``rescore_vis_thr`` is given),
    all float32 or all float64 -- the precision numpy would compute in for the caller's
    records: float32 ndarrays stay float32, Python floats (``.tolist()`` records, what the
    reference's inferencer emits) are float64; image_offset i32 [I+1]; all CUDA tensors.
    Returns (keep i32 [P], num_keep i32 [I]): per image the kept people as local indices in
    keep order, padded with -1."""
    for t, name in ((kpts, "kpts"), (area, "area"), (score, "score"), (image_offset, "image_offset")):
        if not (isinstance(t, torch.Tensor) and t.is_cuda):
            raise ValueError(f"`{name}` must be a CUDA tensor (there is no CPU fallback)")
    if kpts.dtype not in (torch.float32, torch.float64) or area.dtype != kpts.dtype \
            or score.dtype != kpts.dtype:
        raise ValueError("kpts, area and score must be all float32 or all float64")
    if image_offset.dtype != torch.int32:
        raise ValueError("image_offset must be int32")
    if kpts.dim() != 3 or kpts.shape[2] != 3:
        raise ValueError("kpts must be [P, K, 3]")
    if not (kpts.is_contiguous() and area.is_contiguous() and score.is_contiguous()):
        raise ValueError("kpts, area and score must be contiguous")
    people, k = kpts.shape[0], kpts.shape[1]
    num_images = image_offset.numel() - 1
    sig = np.asarray(COCO_SIGMAS if sigmas is None else sigmas, dtype=np.float64)
    if sig.shape != (k,):
        raise ValueError(f"sigmas must have {k} entries")
    p = _lib.OksNmsParams()
    p.num_joints = k
    p.rescore = int(rescore_vis_thr is not None)
    p.use_nms = int(bool(use_nms))
    p.soft = int(bool(soft))
    p.max_dets = int(max_dets)
    p.use_iou_vis_thr = int(iou_vis_thr is not None)
    p.rescore_vis_thr = p.rescore_vis_thr_f64 = \
        float(rescore_vis_thr) if rescore_vis_thr is not None else 0.0
    p.oks_thr = float(oks_thr)
    p.iou_vis_thr = p.iou_vis_thr_f64 = float(iou_vis_thr) if iou_vis_thr is not None else 0.0
    p.max_people_per_image = int(max_people_per_image)
    for j in range(k):
        p.sigmas[j] = float(sig[j])
    dev = kpts.device
    keep = torch.empty(people, dtype=torch.int32, device=dev)
    num_keep = torch.empty(max(num_images, 0), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        entry = "pc_oks_nms_f64" if kpts.dtype == torch.float64 else "pc_oks_nms"
        _lib.call(entry, _lib.device_ptr(kpts), _lib.device_ptr(area),
                  _lib.device_ptr(score), _lib.device_ptr(image_offset), _lib.device_ptr(keep),
                  _lib.device_ptr(num_keep), ctypes.byref(p), num_images, _lib.current_stream())
    return keep, num_keep


def _numpy_precision(*arrays) -> np.dtype:
    """float32 only if numpy itself would stay in float32 for every input (all float32
    ndarrays / scalars); anything else -- Python floats, float64, a mix -- is float64.  (In a
    mix numpy would keep the float32 parts in float32 a little longer; the kernel widens
    first, a difference of at most one float32 ulp in dx, dy before squaring.)"""
    return np.dtype(np.float32) if all(a.dtype == np.float32 for a in arrays) \
        else np.dtype(np.float64)


def _single_image(kpts_db, thr, soft, max_dets, sigmas, vis_thr, device):
    if not kpts_db:
        return []
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    # exactly the arrays the reference builds (nms.py:92-94); their dtype is the precision
    kpts = np.array([np.asarray(k["keypoints"]).flatten() for k in kpts_db])
    area = np.array([k["area"] for k in kpts_db])
    score = np.array([k["score"] for k in kpts_db])
    dt = _numpy_precision(kpts, area, score)
    kpts = np.ascontiguousarray(kpts.reshape(len(kpts_db), -1, 3), dtype=dt)
    area, score = area.astype(dt), score.astype(dt)
    n = len(kpts_db)
    off = torch.tensor([0, n], dtype=torch.int32, device=dev)
    keep, num = rescore_and_nms(torch.from_numpy(kpts).to(dev), torch.from_numpy(area).to(dev),
                                torch.from_numpy(score).to(dev), off, n, oks_thr=thr, soft=soft,
                                max_dets=max_dets, sigmas=sigmas, iou_vis_thr=vis_thr)
    cnt = int(num[0])
    if cnt < 0:
        raise ValueError(f"more than {_lib.PC_NMS_MAX_PEOPLE} people in one image")
    return keep[:cnt].cpu().numpy().astype(np.int64)


def oks_nms(kpts_db: List[Dict[str, Any]], thr: float, sigmas: Optional[np.ndarray] = None,
            vis_thr: Optional[float] = None, device=None) -> np.ndarray:
    """``mindpose.utils.nms.oks_nms`` (nms.py:72-111)."""
    return _single_image(kpts_db, thr, False, 0, sigmas, vis_thr, device)


def soft_oks_nms(kpts_db: List[Dict[str, Any]], thr: float, max_dets: int = 20,
                 sigmas: Optional[np.ndarray] = None, vis_thr: Optional[float] = None,
                 device=None) -> np.ndarray:
    """``mindpose.utils.nms.soft_oks_nms`` (nms.py:141-190)."""
    return _single_image(kpts_db, thr, True, max_dets, sigmas, vis_thr, device)


def evaluate_records(records: Sequence[Dict[str, Any]], config: Dict[str, Any], device=None
                     ) -> List[List[Dict[str, Any]]]:
    """The record handling of ``TopDownEvaluator.eval`` up to the result file
    (topdown_evaluator.py:78-121): group by image (first-seen order), sort and
    de-duplicate by ``bbox_id``, rescore, NMS -- one device launch for all images.
    ``config`` holds ``vis_thr``, ``oks_thr``, ``use_nms``, ``soft_nms``, ``sigmas``.
    Returns, per image, the kept records (copies with the rescored ``score`` / ``area``)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    by_image: Dict[str, List[Dict[str, Any]]] = {}
    for rec in records:
        by_image.setdefault(rec["image_path"].split("/")[-1], []).append(rec)
    groups = []
    for recs in by_image.values():
        recs = sorted(recs, key=lambda r: r["bbox_id"])          # stable, as the reference
        recs = [r for i, r in enumerate(recs) if i == 0 or r["bbox_id"] != recs[i - 1]["bbox_id"]]
        groups.append(recs)
    flat = [r for g in groups for r in g]
    if not flat:
        return []
    counts = [len(g) for g in groups]
    offset = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    # records as the inferencer emits them hold Python lists (`.tolist()`): the reference then
    # computes in float64; float32 ndarrays stay float32 (see _numpy_precision)
    kp_np = np.stack([np.asarray(r["pred"]) for r in flat])
    boxes = np.stack([np.asarray(r["box"]) for r in flat])
    dt = _numpy_precision(kp_np, boxes)
    kpts = torch.from_numpy(np.ascontiguousarray(kp_np, dtype=dt)).to(dev)
    area = torch.from_numpy(np.ascontiguousarray(boxes[:, 4], dtype=dt)).to(dev)
    score = torch.from_numpy(np.ascontiguousarray(boxes[:, 5], dtype=dt)).to(dev)
    keep, num = rescore_and_nms(
        kpts, area, score, torch.from_numpy(offset).to(dev), max(counts),
        oks_thr=config["oks_thr"], rescore_vis_thr=config["vis_thr"],
        use_nms=config.get("use_nms", True), soft=config.get("soft_nms", False),
        sigmas=config.get("sigmas"))
    keep, num, score = keep.cpu().numpy(), num.cpu().numpy(), score.cpu().numpy()
    if (num < 0).any():
        raise ValueError(f"more than {_lib.PC_NMS_MAX_PEOPLE} people in one image")
    out = []
    for g, recs in enumerate(groups):
        kept = []
        for local in keep[offset[g]:offset[g] + num[g]]:
            r = recs[int(local)]
            kept.append(dict(keypoints=r["pred"], center=r["box"][0:2], scale=r["box"][2:4],
                             area=r["box"][4], score=score[offset[g] + int(local)],
                             bbox_id=r["bbox_id"], image_path=r["image_path"]))
        out.append(kept)
    return out
