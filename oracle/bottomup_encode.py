"""Oracle (test infrastructure): bottom-up multi-resolution target encoding.

Restates ``BottomUpGenerateTarget._encoding`` / ``._generate_heatmap_and_tag_ind``
(mindpose/data/transform/bottomup_transform.py:504-598) and ``pad_to_same``
(mindpose/data/transform/utils.py:213-232) in numpy -- SURVEY.md section 8(f) row N1.

Organised per (person, joint) as "window geometry, then values"; every arithmetic
step keeps the reference's dtype under numpy >= 2 (NEP 50: Python scalars are weak):

* the centre is Python ``round`` of the float32 coordinate (half to even, :560);
* the window is ``[mu - 3 sigma, mu + 3 sigma + 1)`` with ``int()`` truncation (:563-564),
  skipped when it lies fully outside the map (:565-566);
* the Gaussian is centred at the sub-pixel position ``x0 + pt - mu`` -- float32, because
  ``x0`` (a Python float) and ``mu`` (a Python int) are weak scalars -- and evaluated in
  float32: ``exp(-((x - x0_p)**2 + (y - y0_p)**2) / (2 sigma**2))`` (:568-572);
* overlapping windows are merged with ``np.maximum`` (:585-588), so the result does not
  depend on the order of the people;
* ``tag_ind[m, k] = (mu_y * W + mu_x, 1)`` when the centre itself is inside the map
  (:590-596), zeros otherwise; with ``tag_per_joint=False`` the last visible joint of a
  person wins;
* the per-scale maps are zero-padded at the bottom / right to the largest one (:519-522).

PINNED by tests/golden/bottomup_encode_ref.npz (outputs of the imported reference).
"""
import numpy as np


def generate_heatmap_and_tag_ind(keypoints, heatmap_size, sigma=2.0, max_num=30,
                                 tag_per_joint=True):
    """keypoints float32 [M, K, 3] in heat-map pixels; heatmap_size = (W, H).
    -> (target float32 [K, H, W], tag_ind int32 [max_num, K, 2] or [max_num, 2])."""
    w, h = int(heatmap_size[0]), int(heatmap_size[1])
    keypoints = np.asarray(keypoints)
    m_people, k_joints, _ = keypoints.shape
    if m_people > max_num:
        raise ValueError(
            f"Number of keypoints in one image `{m_people}` exeeds the maximum num: `{max_num}`")
    target = np.zeros((k_joints, h, w), np.float32)
    tag_ind = np.zeros((max_num, k_joints, 2) if tag_per_joint else (max_num, 2), np.int32)
    tmp = sigma * 3
    size = 2 * tmp + 1
    grid = np.arange(0, size, 1, np.float32)
    c0 = size // 2
    two_sigma2 = 2 * sigma ** 2
    for m in range(m_people):
        for k in range(k_joints):
            px, py, vis = keypoints[m, k]
            if not vis > 0:
                continue
            mu_x, mu_y = round(px), round(py)
            ul_x, ul_y = int(mu_x - tmp), int(mu_y - tmp)
            br_x, br_y = int(mu_x + tmp + 1), int(mu_y + tmp + 1)
            if ul_x >= w or ul_y >= h or br_x < 0 or br_y < 0:
                continue
            cx = c0 + px - mu_x  # float32 under NEP 50
            cy = c0 + py - mu_y
            x_lo, x_hi = max(0, ul_x), min(br_x, w)
            y_lo, y_hi = max(0, ul_y), min(br_y, h)
            gx = (grid[x_lo - ul_x:x_hi - ul_x] - cx) ** 2
            gy = (grid[y_lo - ul_y:y_hi - ul_y] - cy) ** 2
            g = np.exp(-(gx[None, :] + gy[:, None]) / two_sigma2)
            patch = target[k, y_lo:y_hi, x_lo:x_hi]
            target[k, y_lo:y_hi, x_lo:x_hi] = np.maximum(patch, g)
            if mu_x >= w or mu_y >= h or mu_x < 0 or mu_y < 0:
                continue
            if tag_per_joint:
                tag_ind[m, k] = (mu_y * w + mu_x, 1)
            else:
                tag_ind[m] = (mu_y * w + mu_x, 1)
    return target, tag_ind


def encode(keypoints_per_scale, heatmap_sizes, sigma=2.0, max_num=30, tag_per_joint=True):
    """``_encoding``: one keypoint array per scale -> (target [S,K,Hmax,Wmax], tag_ind [S,...])."""
    targets, tags = [], []
    for kps, size in zip(keypoints_per_scale, heatmap_sizes):
        t, g = generate_heatmap_and_tag_ind(kps, size, sigma, max_num, tag_per_joint)
        targets.append(t)
        tags.append(g)
    hmax = max(t.shape[1] for t in targets)
    wmax = max(t.shape[2] for t in targets)
    out = np.zeros((len(targets), targets[0].shape[0], hmax, wmax), np.float32)
    for s, t in enumerate(targets):
        out[s, :, :t.shape[1], :t.shape[2]] = t
    return out, np.stack(tags)
