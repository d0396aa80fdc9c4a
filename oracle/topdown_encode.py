"""Oracle (test infrastructure): top-down heatmap target encoding.

Restates ``TopDownGenerateTarget._encoding`` and ``._udp_encoding``
(mindpose/data/transform/topdown_transform.py:324-375 and :377-430) in numpy.
The restatement is organised per joint as "window geometry, then values" rather
than as the reference's paste of a pre-built patch, but every arithmetic step
keeps the reference's dtype:

* feat_stride is float64 (``image_size / [W, H]`` resp.
  ``(image_size - 1) / ([W, H] - 1)``), so ``kp / feat_stride`` is a float64
  quotient of a float32 keypoint;
* the standard centre is Python ``round`` (half to even, :350-351), the UDP
  centre is ``int(q + 0.5)`` (truncation toward zero, :399-400);
* the standard 13x13 patch is evaluated in float32 (:339-344); the UDP patch is
  evaluated around the sub-pixel centre in float64 and rounded to float32 when
  stored (numpy >= 2 promotion; numpy 1.x evaluates it in float32, <= 1e-7
  apart, inside the 1e-5 tolerance of the north star);
* ``target_weight = visibility``, zeroed when the window is fully outside
  (:353-357); the patch is written only when the weight exceeds 0.5 (:359).

PINNED by tests/golden/encode_*.npz (outputs of the imported reference).
"""
import numpy as np


def _window(mu_x, mu_y, tmp_size, w, h):
    ul = [int(mu_x - tmp_size), int(mu_y - tmp_size)]
    br = [int(mu_x + tmp_size + 1), int(mu_y + tmp_size + 1)]
    outside = ul[0] >= w or ul[1] >= h or br[0] < 0 or br[1] < 0
    return ul, br, outside


def encode_gaussian(keypoints, image_size, heatmap_size, sigma=2.0, joint_weights=None):
    """keypoints f32 [K, 3]; image_size/heatmap_size = [w, h].

    Returns (target f32 [K, H, W], target_weight f32 [K])."""
    image_size = np.asarray(image_size)
    w, h = int(heatmap_size[0]), int(heatmap_size[1])
    k = keypoints.shape[0]
    target = np.zeros((k, h, w), dtype=np.float32)
    weight = np.zeros(k, dtype=np.float32)
    tmp_size = sigma * 3
    size = 2 * tmp_size + 1
    c0 = size // 2
    t = np.arange(0, size, 1, np.float32)
    patch = np.exp(-((t[None, :] - c0) ** 2 + (t[:, None] - c0) ** 2) / (2 * sigma**2))
    feat_stride = image_size / np.array([w, h])
    for j in range(k):
        weight[j] = keypoints[j, 2]
        mu_x = round(keypoints[j][0] / feat_stride[0])
        mu_y = round(keypoints[j][1] / feat_stride[1])
        ul, br, outside = _window(mu_x, mu_y, tmp_size, w, h)
        if outside:
            weight[j] = 0
            continue
        if weight[j] > 0.5:
            x_lo, x_hi = max(0, ul[0]), min(br[0], w)
            y_lo, y_hi = max(0, ul[1]), min(br[1], h)
            target[j, y_lo:y_hi, x_lo:x_hi] = patch[
                y_lo - ul[1] : y_hi - ul[1], x_lo - ul[0] : x_hi - ul[0]
            ]
    if joint_weights is not None:
        weight = np.multiply(weight, joint_weights)
    return target, weight


def encode_udp(keypoints, image_size, heatmap_size, sigma=2.0, joint_weights=None):
    """UDP variant: sub-pixel centred Gaussian, heatmap only (the reference has
    no offset-map encoding)."""
    image_size = np.asarray(image_size)
    w, h = int(heatmap_size[0]), int(heatmap_size[1])
    k = keypoints.shape[0]
    target = np.zeros((k, h, w), dtype=np.float32)
    weight = np.zeros(k, dtype=np.float32)
    tmp_size = sigma * 3
    size = 2 * tmp_size + 1
    c0 = size // 2
    t = np.arange(0, size, 1, np.float32)
    feat_stride = (image_size - 1.0) / (np.array([w, h]) - 1.0)
    for j in range(k):
        weight[j] = keypoints[j, 2]
        qx = keypoints[j][0] / feat_stride[0]
        qy = keypoints[j][1] / feat_stride[1]
        mu_x = int(qx + 0.5)
        mu_y = int(qy + 0.5)
        ul, br, outside = _window(mu_x, mu_y, tmp_size, w, h)
        if outside:
            weight[j] = 0
            continue
        if weight[j] > 0.5:
            cx = c0 + qx - mu_x
            cy = c0 + qy - mu_y
            patch = np.exp(
                -((t[None, :] - cx) ** 2 + (t[:, None] - cy) ** 2) / (2 * sigma**2)
            )
            x_lo, x_hi = max(0, ul[0]), min(br[0], w)
            y_lo, y_hi = max(0, ul[1]), min(br[1], h)
            target[j, y_lo:y_hi, x_lo:x_hi] = patch[
                y_lo - ul[1] : y_hi - ul[1], x_lo - ul[0] : x_hi - ul[0]
            ]
    if joint_weights is not None:
        weight = np.multiply(weight, joint_weights)
    return target, weight
