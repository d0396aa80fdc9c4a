"""Generate tests/golden/*.npz (run in the build container only).

    python -m oracle.gen_golden

* ``*_ref.npz`` are outputs of the UNMODIFIED reference functions imported from
  /root/reference (oracle/ref_loader.py), or of the third-party libraries the
  reference calls (cv2.warpAffine, scipy linear_sum_assignment).  They pin the
  numpy restatements in this package.
* ``*_restated.npz`` are outputs of the restatements of the MindSpore half
  (parity unpinned); they are frozen so the oracle of record cannot drift
  unnoticed between rounds.

The inputs are stored next to the outputs, so the tests never need the
reference tree.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from mindpose_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402


def _cfg(ns, image_size, heatmap_size, **extra):
    cfg = dict(synth.TOPDOWN_CONFIG, image_size=image_size, heatmap_size=heatmap_size)
    cfg.update(extra)
    return cfg


def gen_affine(ns):
    rng = np.random.RandomState(11)
    n = 64
    boxes = np.stack(
        [rng.uniform(-20, 500, n), rng.uniform(-20, 400, n), rng.uniform(5, 400, n),
         rng.uniform(5, 440, n)], axis=1).astype(np.float32)
    boxes[0] = [10, 20, 192, 256]  # exact aspect ratio: neither branch
    boxes[1] = [0, 0, 96, 128]
    rots = rng.uniform(-80, 80, n)
    rots[: n // 2] = 0.0
    out = dict(boxes=boxes, rots=rots)
    for tag, image_size in (("256x192", [192, 256]), ("384x288", [288, 384])):
        cfg = _cfg(ns, image_size, [image_size[0] // 4, image_size[1] // 4])
        t = ns.topdown.TopDownBoxToCenterScale(is_train=False, config=cfg)
        cs = [t._xywh2cs(*b) for b in boxes]
        center = np.stack([c for c, _ in cs])
        scale = np.stack([s for _, s in cs])
        std = np.stack([
            ns.utils.get_affine_transform(c, s, float(r), np.array(image_size), pixel_std=200.0)
            for c, s, r in zip(center, scale, rots)])
        udp = np.stack([
            ns.utils.get_warp_matrix(float(r), c * 2.0, np.array(image_size) - 1.0, s * 200.0)
            for c, s, r in zip(center, scale, rots)])
        kps = synth.keypoints(n, 17, [640, 480], seed=5)
        kp_std, kp_udp = [], []
        for i in range(n):
            aff = ns.topdown.TopDownAffine(is_train=False, config=cfg)
            k1 = kps[i].copy()
            for j in range(17):
                if k1[j, 2] > 0.0:
                    k1[j, 0:2] = ns.utils.affine_transform(k1[j, 0:2], std[i])
            kp_std.append(k1)
            k2 = kps[i].copy()
            k2[:, 0:2] = ns.utils.warp_affine_joints(k2[:, 0:2], udp[i])
            kp_udp.append(k2)
            del aff
        out.update({
            f"center_{tag}": center, f"scale_{tag}": scale, f"std_{tag}": std,
            f"udp_{tag}": udp, f"kps_in": kps, f"kps_std_{tag}": np.stack(kp_std),
            f"kps_udp_{tag}": np.stack(kp_udp)})
    np.savez_compressed(os.path.join(GOLDEN, "affine_ref.npz"), **out)


def gen_affine_rotated(ns):
    """The reference's own ``TopDownAffine.transform`` (matrix + cv2.warpAffine + joints) on
    rotated samples, fed the way its dataset pipeline feeds it: float32 arrays and a float32
    0-d rotation (transform.py:66-79).  -> affine_rot_ref.npz."""
    rng = np.random.RandomState(23)
    n, hs, ws = 12, 120, 160
    image_size = [96, 128]
    images = rng.randint(0, 256, size=(n, hs, ws, 3), dtype=np.uint8)
    boxes = np.stack([rng.uniform(0, 60, n), rng.uniform(0, 40, n), rng.uniform(30, 100, n),
                      rng.uniform(40, 80, n)], axis=1).astype(np.float32)
    rots = rng.uniform(-60, 60, n).astype(np.float32)
    rots[0] = 0.0
    rots[1] = 90.0
    kps = synth.keypoints(n, 17, [ws, hs], seed=8)
    out = dict(images=images, boxes=boxes, rots=rots, kps_in=kps,
               image_size=np.asarray(image_size))
    for tag, use_udp in (("std", False), ("udp", True)):
        cfg = _cfg(ns, np.array(image_size), np.array([image_size[0] // 4, image_size[1] // 4]))
        bt = ns.topdown.TopDownBoxToCenterScale(is_train=False, config=cfg)
        at = ns.topdown.TopDownAffine(is_train=False, config=cfg, use_udp=use_udp)
        crops, kout, centers, scales, mats = [], [], [], [], []
        for i in range(n):
            c, sc = bt._xywh2cs(*boxes[i])
            r0 = np.asarray(rots[i])
            mats.append(ns.utils.get_warp_matrix(r0, c * 2.0, np.array(image_size) - 1.0, sc * 200.0)
                        if use_udp else
                        ns.utils.get_affine_transform(c, sc, r0, np.array(image_size), pixel_std=200.0))
            state = dict(image=images[i], center=np.asarray(c), scale=np.asarray(sc),
                         rotation=np.asarray(rots[i]), keypoints=kps[i].copy())
            res = at.transform(state)
            crops.append(res["image"])
            kout.append(np.asarray(res["keypoints"]))
            centers.append(c)
            scales.append(sc)
        out.update({f"crops_{tag}": np.stack(crops), f"kps_{tag}": np.stack(kout),
                    f"mats_{tag}": np.stack(mats).astype(np.float64),
                    "center": np.stack(centers), "scale": np.stack(scales)})
    np.savez_compressed(os.path.join(GOLDEN, "affine_rot_ref.npz"), **out)


def gen_encode(ns):
    out = {}
    for tag, image_size, heatmap_size, n in (("64x48", [192, 256], [48, 64], 12),
                                            ("96x72", [288, 384], [72, 96], 6)):
        kps = synth.keypoints(n, 17, image_size, seed=3)
        # exact half-way cases for round() / int(x + 0.5), and a visible joint far outside
        kps[0, 0, :2] = [10.0, 30.0]
        kps[0, 1, :2] = [6.0, 14.0]
        kps[0, 2, :2] = [-2.0, -6.0]
        kps[0, 3] = [-100.0, 50.0, 1.0]
        kps[0, 4] = [image_size[0] + 27.9, 50.0, 1.0]
        kps[0, 5] = [-28.0, -28.0, 1.0]
        kps[0, 6] = [-27.9, 40.0, 1.0]
        kps[0, 7, 2] = 2.0
        kps[0, 8, 2] = 0.5
        cfg = _cfg(ns, image_size, heatmap_size)
        out[f"kps_{tag}"] = kps
        for udp in (False, True):
            t = ns.topdown.TopDownGenerateTarget(is_train=True, config=cfg, sigma=2.0, use_udp=udp)
            res = [t.transform(dict(keypoints=k.copy())) for k in kps]
            name = "udp" if udp else "std"
            out[f"target_{name}_{tag}"] = np.stack([r["target"] for r in res])
            out[f"weight_{name}_{tag}"] = np.stack([r["target_weight"] for r in res])
    # sigma = 3 and joint weights
    jw = np.linspace(1.0, 1.5, 17)
    cfg = _cfg(ns, [192, 256], [48, 64], joint_weights=jw.tolist())
    t = ns.topdown.TopDownGenerateTarget(is_train=True, config=cfg, sigma=3.0,
                                        use_different_joint_weights=True)
    kps = out["kps_64x48"][:4]
    res = [t.transform(dict(keypoints=k.copy())) for k in kps]
    out["joint_weights"] = jw
    out["target_std_sigma3"] = np.stack([r["target"] for r in res])
    out["weight_std_sigma3"] = np.stack([r["target_weight"] for r in res])
    np.savez_compressed(os.path.join(GOLDEN, "encode_ref.npz"), **out)


def gen_warp(ns):
    import cv2

    rng = np.random.RandomState(21)
    srcs, mats, outs, sizes = [], [], [], []
    for i in range(10):
        hs, ws = (90, 120) if i % 2 == 0 else (75, 133)
        img = rng.randint(0, 256, (hs, ws, 3)).astype(np.uint8)
        box = (np.float32(rng.uniform(-10, ws * 0.5)), np.float32(rng.uniform(-10, hs * 0.5)),
               np.float32(rng.uniform(15, ws)), np.float32(rng.uniform(15, hs)))
        dsize = (48, 64) if i < 6 else (72, 96)
        cfg = _cfg(ns, list(dsize), [dsize[0] // 4, dsize[1] // 4])
        t = ns.topdown.TopDownBoxToCenterScale(is_train=False, config=cfg)
        c, s = t._xywh2cs(*box)
        rot = 0.0 if i % 3 == 0 else float(rng.uniform(-60, 60))
        if i % 2 == 0:
            m = ns.utils.get_affine_transform(c, s, rot, np.array(dsize), pixel_std=200.0)
        else:
            m = ns.utils.get_warp_matrix(rot, c * 2.0, np.array(dsize) - 1.0, s * 200.0)
        dst = cv2.warpAffine(img, m, dsize, flags=cv2.INTER_LINEAR)
        srcs.append(img.reshape(-1))
        sizes.append([hs, ws, dsize[0], dsize[1]])
        mats.append(np.asarray(m, dtype=np.float64))
        outs.append(dst.reshape(-1))
    np.savez_compressed(
        os.path.join(GOLDEN, "warp_ref.npz"),
        src=np.concatenate(srcs), dst=np.concatenate(outs), sizes=np.array(sizes),
        mats=np.stack(mats), cv2_version=np.array(cv2.__version__))


def gen_decode_restated():
    from oracle import topdown_decode as td

    fidx = synth.flip_index()
    out = {}
    for tag, h, w in (("64x48", 64, 48), ("96x72", 96, 72)):
        n = 6
        blobs, _ = synth.blob_heatmaps(n, 17, h, w, seed=1)
        flipped = synth.flipped_pair(blobs, seed=1)
        center, scale, score = synth.crop_geometry(n, seed=1)
        out[f"center_{tag}"], out[f"scale_{tag}"], out[f"score_{tag}"] = center, scale, score
        p, b = td.decode(blobs, center, scale, score)
        out[f"plain_preds_{tag}"], out[f"plain_boxes_{tag}"] = p, b
        p, _ = td.decode(blobs, center, scale, score, shift_coordinate_flag=True)
        out[f"shift_preds_{tag}"] = p
        p, _ = td.decode(blobs, center, scale, score, dark_udp_refine_flag=True, use_udp=True)
        out[f"dark_udp_preds_{tag}"] = p
        p, _ = td.decode_with_flip(blobs, flipped, fidx, center, scale, score,
                                   shift_heatmap=True, shift_coordinate_flag=True)
        out[f"flip_shift_preds_{tag}"] = p
        p, _ = td.decode_with_flip(blobs, flipped, fidx, center, scale, score,
                                   shift_heatmap=False, dark_udp_refine_flag=True)
        out[f"flip_dark_preds_{tag}"] = p
    np.savez_compressed(os.path.join(GOLDEN, "topdown_decode_restated.npz"), **out)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ns = ref_loader.load()
    gen_affine(ns)
    gen_affine_rotated(ns)
    gen_encode(ns)
    gen_warp(ns)
    gen_decode_restated()
    try:
        from oracle import gen_golden_bottomup

        gen_golden_bottomup.main(ns, GOLDEN)
    except ImportError:
        pass
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
