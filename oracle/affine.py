"""Oracle (test infrastructure): crop geometry of the top-down codec.

Restates, in numpy, the scalar geometry the reference does on the host:

* ``box_to_center_scale``  -- ``TopDownBoxToCenterScale._xywh2cs``
  (mindpose/data/transform/topdown_transform.py:131-154, eval branch: no
  random centre shift).
* ``affine_matrix``        -- ``get_affine_transform``
  (mindpose/data/transform/utils.py:44-98, with ``rotate_point`` :117 and
  ``_get_3rd_point`` :136).  The three source / destination points are stored
  in float32 exactly as the reference does (:83-91); the 3-point solve that the
  reference delegates to ``cv2.getAffineTransform`` (third-party) is restated
  op for op (6x6 LU with partial pivoting, fp64).
* ``udp_matrix``           -- ``get_warp_matrix`` (utils.py:158-190), float32
  result, as called by ``TopDownAffine._udp_affine``
  (topdown_transform.py:239-244).
* ``transform_joints`` / ``transform_joints_udp`` -- the keypoint half of
  ``_affine`` (:224-231, only joints with visibility > 0) and ``_udp_affine``
  (:255-259, all joints).

PINNED by tests/golden/affine_*.npz (outputs of the imported reference).
"""
import numpy as np


def box_to_center_scale(box, image_size, pixel_std=200.0, scale_padding=1.25):
    """box = (x, y, w, h); image_size = [w, h]. Returns (center f32[2], scale f32[2])."""
    x, y, w, h = box
    aspect = image_size[0] / image_size[1]
    center = np.array([x + w * 0.5, y + h * 0.5], dtype=np.float32)
    if w > aspect * h:
        h = w * 1.0 / aspect
    elif w < aspect * h:
        w = h * aspect
    scale = np.array([w / pixel_std, h / pixel_std], dtype=np.float32)
    scale = scale * scale_padding
    return center, scale


def _third_point(a, b):
    d = a - b
    return b + np.array([-d[1], d[0]], dtype=np.float32)


def _solve_three_points(src, dst):
    """``cv2.getAffineTransform`` restated: the 6x6 system
    ``[x y 1 0 0 0; 0 0 0 x y 1] m = d`` solved by OpenCV's in-house LU with partial
    pivoting in fp64 (third-party algorithm, modules/core matrix_decomp ``LUImpl``).
    Bit-identical to cv2 4.13.0 on the golden matrices; the last bit matters
    because the warp's fixed-point rounding frequently sits on an exact tie."""
    a = np.zeros((6, 6), dtype=np.float64)
    b = np.zeros(6, dtype=np.float64)
    for i in range(3):
        a[2 * i, 0] = a[2 * i + 1, 3] = src[i, 0]
        a[2 * i, 1] = a[2 * i + 1, 4] = src[i, 1]
        a[2 * i, 2] = a[2 * i + 1, 5] = 1.0
        b[2 * i] = dst[i, 0]
        b[2 * i + 1] = dst[i, 1]
    m = 6
    for i in range(m):
        k = i
        for j in range(i + 1, m):
            if abs(a[j, i]) > abs(a[k, i]):
                k = j
        if k != i:
            a[[i, k], i:] = a[[k, i], i:]
            b[[i, k]] = b[[k, i]]
        d = -1.0 / a[i, i]
        for j in range(i + 1, m):
            alpha = a[j, i] * d
            for c in range(i + 1, m):
                a[j, c] = a[j, c] + alpha * a[i, c]
            b[j] = b[j] + alpha * b[i]
    for i in range(m - 1, -1, -1):
        s = b[i]
        for c in range(i + 1, m):
            s = s - a[i, c] * b[c]
        b[i] = s / a[i, i]
    return b.reshape(2, 3).copy()


def affine_matrix(center, scale, rot, output_size, pixel_std=200.0, inv=False):
    """Standard (non-UDP) crop matrix, float64 [2, 3]."""
    center = np.asarray(center)
    scale = np.asarray(scale)
    scale_tmp = scale * pixel_std
    src_w = scale_tmp[0]
    dst_w = output_size[0]
    dst_h = output_size[1]
    rot_rad = np.pi * rot / 180
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    px, py = 0.0, src_w * -0.5
    src_dir = [px * cs - py * sn, px * sn + py * cs]
    dst_dir = np.array([0.0, dst_w * -0.5])

    src = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center + scale_tmp * np.array([0.0, 0.0])
    src[1, :] = center + src_dir + scale_tmp * np.array([0.0, 0.0])
    src[2, :] = _third_point(src[0, :], src[1, :])
    dst = np.zeros((3, 2), dtype=np.float32)
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir
    dst[2, :] = _third_point(dst[0, :], dst[1, :])
    if inv:
        return _solve_three_points(dst, src)
    return _solve_three_points(src, dst)


def udp_matrix(center, scale, rot, image_size, pixel_std=200.0):
    """UDP crop matrix, float32 [2, 3] (reference stores it in float32)."""
    center = np.asarray(center)
    scale = np.asarray(scale)
    size_input = center * 2.0
    size_dst = np.asarray(image_size) - 1.0
    size_target = scale * pixel_std
    theta = np.deg2rad(rot)
    m = np.zeros((2, 3), dtype=np.float32)
    sx = size_dst[0] / size_target[0]
    sy = size_dst[1] / size_target[1]
    m[0, 0] = np.cos(theta) * sx
    m[0, 1] = -np.sin(theta) * sx
    m[0, 2] = sx * (
        -0.5 * size_input[0] * np.cos(theta)
        + 0.5 * size_input[1] * np.sin(theta)
        + 0.5 * size_target[0]
    )
    m[1, 0] = np.sin(theta) * sy
    m[1, 1] = np.cos(theta) * sy
    m[1, 2] = sy * (
        -0.5 * size_input[0] * np.sin(theta)
        - 0.5 * size_input[1] * np.cos(theta)
        + 0.5 * size_target[1]
    )
    return m


def transform_joints(keypoints, mat):
    """Standard path: joints with visibility > 0 go through mat (fp64 matmul,
    stored back into the keypoints' own dtype)."""
    out = keypoints.copy()
    for i in range(out.shape[0]):
        if out[i, 2] > 0.0:
            out[i, 0:2] = np.array(mat) @ np.array([out[i, 0], out[i, 1], 1.0])
    return out


def transform_joints_udp(keypoints, mat):
    """UDP path: all joints, ``[x, y, 1] @ mat.T``."""
    out = keypoints.copy()
    xy1 = np.concatenate(
        (out[:, 0:2], np.ones((out.shape[0], 1), dtype=np.float32)), axis=-1
    )
    out[:, 0:2] = np.dot(xy1, mat.T)
    return out
