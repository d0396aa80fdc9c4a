"""Oracle (test infrastructure): associative-embedding tag grouping.

Restates ``match_by_tag`` (mindpose/utils/match.py:14-116), the instance score of
``BottomUpHeatMapAEInferencer._parse``
(mindpose/engine/inferencer/bottomup_inferencer.py:153-156) and
``transform_keypoints`` (mindpose/data/transform/utils.py:235-274).

The reference keeps groups in two dicts keyed by the float tag value of the
detection that opened the group.  The restatement keeps an explicit, insertion
ordered list of groups and reproduces what the dict semantics imply:

* a NEW detection whose first tag value equals (==) the key of an existing
  group does not open a new group: it overwrites that group's row for the
  current joint and RESETS the group's tag list to this one tag (:65-67,
  :110-113); the group keeps its place in the order;
* ``-0.0`` and ``0.0`` are the same key (they hash and compare equal);
* group reference tag = float32 mean of the tag list (numpy pairwise sum, which
  for <= 17 entries of one float is: plain left-to-right below 8 entries, eight
  strided partial sums combined pairwise from 8 up);
* distance = sqrt(sum(diff^2)) in float32, rounded half-to-even when
  ``use_rounded_norm``; padded with 1e10 columns to a square when there are more
  detections than groups; assignment by ``oracle.lsap`` (scipy's algorithm);
* a pair is accepted iff the UNROUNDED distance < tag_thr.

PINNED by tests/golden/match_ref.npz (outputs of the imported reference).
"""
import numpy as np

from . import lsap

F32 = np.float32


def _pairwise_sum_f32(vals):
    """numpy's float32 add.reduce over a short contiguous run."""
    n = len(vals)
    if n < 8:
        res = F32(0.0)
        for v in vals:
            res = F32(res + v)
        return res
    r = [F32(vals[i]) for i in range(8)]
    i = 8
    while i < n - (n % 8):
        for t in range(8):
            r[t] = F32(r[t] + vals[i + t])
        i += 8
    res = F32(F32(F32(r[0] + r[1]) + F32(r[2] + r[3])) + F32(F32(r[4] + r[5]) + F32(r[6] + r[7])))
    while i < n:
        res = F32(res + vals[i])
        i += 1
    return res


def _mean_f32(vecs):
    """np.mean(np.stack(list_of_[L]_float32), axis=0) -> float32 [L]."""
    n = len(vecs)
    length = len(vecs[0])
    out = np.zeros(length, dtype=F32)
    for c in range(length):
        out[c] = F32(_pairwise_sum_f32([v[c] for v in vecs]) / F32(n))
    return out


def match_by_tag(val_k, tag_k, ind_k, joint_order, vis_thr=0.1, tag_thr=1.0,
                 ignore_too_much=False, use_rounded_norm=True):
    """val_k [K,M], tag_k [K,M,L], ind_k [K,M,2] -> float32 [P,K,3+L] (empty: shape (0,))."""
    val_k = np.asarray(val_k, dtype=F32)
    tag_k = np.asarray(tag_k, dtype=F32)
    ind_k = np.asarray(ind_k, dtype=F32)
    num_joints, max_num, tag_len = tag_k.shape
    joint_k = np.concatenate((ind_k, val_k[..., None], tag_k), axis=2)

    keys = []    # float32 key per group, insertion order
    rows = []    # [K, 3+L] per group
    tags = []    # list of tag vectors per group

    def find(key):
        for g, kv in enumerate(keys):
            if kv == key:
                return g
        return -1

    def open_or_overwrite(key, idx, joint, tag):
        g = find(key)
        if g < 0:
            keys.append(key)
            rows.append(np.zeros((num_joints, 3 + tag_len), dtype=F32))
            tags.append(None)
            g = len(keys) - 1
        rows[g][idx] = joint
        tags[g] = [tag]

    for i in range(num_joints):
        idx = joint_order[i]
        keep = joint_k[idx][:, 2] > F32(vis_thr)
        tg = tag_k[idx][keep]
        if tg.shape[0] == 0:
            continue
        jt = joint_k[idx][keep]
        if i == 0 or len(keys) == 0:
            for j in range(tg.shape[0]):
                open_or_overwrite(tg[j, 0], idx, jt[j], tg[j])
            continue
        num_grouped = len(keys)
        if ignore_too_much and num_grouped == max_num:
            continue
        ref_tags = np.stack([_mean_f32(t) for t in tags])          # [G, L]
        diff = jt[:, None, 3:] - ref_tags[None, :, :]
        sq = (diff * diff).astype(F32)
        acc = np.zeros(sq.shape[:2], dtype=F32)
        for c in range(tag_len):
            acc = (acc + sq[:, :, c]).astype(F32)
        dist = np.sqrt(acc).astype(F32)
        saved = dist.copy()
        if use_rounded_norm:
            dist = np.round(dist)
        num_added = dist.shape[0]
        if num_added > num_grouped:
            pad = np.zeros((num_added, num_added - num_grouped), F32) + F32(1e10)
            dist = np.concatenate((dist, pad), axis=1)
        rr, cc = lsap.linear_sum_assignment(dist)
        for row, col in zip(rr, cc):
            if row < num_added and col < num_grouped and saved[row][col] < tag_thr:
                rows[col][idx] = jt[row]
                tags[col].append(tg[row])
            else:
                open_or_overwrite(tg[row, 0], idx, jt[row], tg[row])
    if not rows:
        return np.zeros((0,), dtype=F32)
    return np.stack(rows).astype(F32)


def instance_scores(ans):
    """Mean of the value column over all joints (missing joints count as 0)."""
    return [F32(_pairwise_sum_f32(list(p[:, 2])) / F32(p.shape[0])) for p in ans]


def transform_keypoints(coords, center, scale, heatmap_shape, pixel_std=200.0):
    """coords: list of [P,K,>=2] arrays (one per image); center/scale/heatmap_shape [N,2]."""
    scale = scale * pixel_std
    sx = scale[:, 0] / heatmap_shape[:, 0]
    sy = scale[:, 1] / heatmap_shape[:, 1]
    out = []
    for i, c in enumerate(coords):
        if c.size == 0:
            out.append(c)
            continue
        t = c.copy()
        t[:, :, 0] = c[:, :, 0] * sx[i] + center[i, 0] - scale[i, 0] * 0.5
        t[:, :, 1] = c[:, :, 1] * sy[i] + center[i, 1] - scale[i, 1] * 0.5
        out.append(t)
    return out
