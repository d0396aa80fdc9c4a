"""Bottom-up half of oracle/gen_golden.py (see there)."""
import os

import numpy as np

from mindpose_b200 import synth


def grouping_inputs(seed, n, k=17, m=30, mode="people"):
    """Synthetic (val_k, tag_k, ind_k) batches that exercise match_by_tag."""
    rng = np.random.RandomState(seed)
    val = np.zeros((n, k, m), np.float32)
    tag = np.zeros((n, k, m, 1), np.float32)
    ind = np.zeros((n, k, m, 2), np.float32)
    for i in range(n):
        people = rng.randint(0, 14)
        for j in range(k):
            ids = rng.permutation(people)
            ndet = 0
            for pid in ids:
                if rng.random_sample() < 0.2 or ndet >= m:
                    continue
                if mode == "people":
                    t = pid * 3.0 + rng.normal(0, 0.15)
                elif mode == "ties":      # integer tags: rounded distances tie everywhere
                    t = float(pid % 5) + (0.0 if rng.random_sample() < 0.7 else 0.5)
                else:                      # "crowded": tags close together
                    t = pid * 0.6 + rng.normal(0, 0.2)
                val[i, j, ndet] = rng.uniform(0.05, 1.0)
                tag[i, j, ndet, 0] = t
                ind[i, j, ndet] = rng.randint(0, 256, 2)
                ndet += 1
            # spurious low / high detections
            extra = rng.randint(0, 4)
            for _ in range(extra):
                if ndet >= m:
                    break
                val[i, j, ndet] = rng.uniform(0.0, 0.5)
                tag[i, j, ndet, 0] = rng.uniform(-1, 40) if mode != "ties" else float(rng.randint(0, 6))
                ind[i, j, ndet] = rng.randint(0, 256, 2)
                ndet += 1
            order = np.argsort(-val[i, j], kind="stable")
            val[i, j] = val[i, j][order]
            tag[i, j] = tag[i, j][order]
            ind[i, j] = ind[i, j][order]
    return val, tag, ind


BOTTOMUP_ENCODE_CASES = [
    # (heatmap_sizes [[w, h], ...], people, tag_per_joint, seed)
    ([[32, 32], [64, 64]], 0, True, 0),
    ([[32, 32], [64, 64]], 1, True, 1),
    ([[32, 32], [64, 64]], 8, True, 2),
    ([[32, 32], [64, 64]], 30, True, 3),
    ([[48, 40], [96, 80]], 12, True, 4),
    ([[32, 32], [64, 64]], 8, False, 5),
    ([[64, 64]], 5, True, 6),
]


def bottomup_people(seed, m, k, sizes):
    """One keypoint array per scale (float32 [M,K,3], heat-map pixels of that scale): windows
    that clip, fall outside, overlap (np.maximum merge) and exact .5 coordinates."""
    rng = np.random.RandomState(seed)
    w0, h0 = sizes[0]
    kp = np.zeros((m, k, 3), np.float32)
    kp[..., 0] = rng.uniform(-10, w0 + 10, (m, k))
    kp[..., 1] = rng.uniform(-10, h0 + 10, (m, k))
    kp[..., 2] = rng.choice([0, 1, 2], (m, k), p=[0.2, 0.3, 0.5])
    if m > 0:
        kp[0, :, 0] = np.floor(kp[0, :, 0]) + 0.5
    if m > 2:
        kp[2, :, :2] = kp[1, :, :2] + rng.uniform(-3, 3, (k, 2))
    out = []
    for (w, h) in sizes:
        s = kp.copy()
        s[..., 0] *= np.float32(w / w0)
        s[..., 1] *= np.float32(h / h0)
        out.append(s)
    return out


def bottomup_encode_golden(ns, golden_dir):
    out = {}
    for ci, (sizes, m, tpj, seed) in enumerate(BOTTOMUP_ENCODE_CASES):
        cfg = dict(image_size=[512, 512], max_image_size=[832, 512], heatmap_sizes=sizes,
                   flip_pairs=[[1, 2]], pixel_std=200.0, tag_per_joint=tpj)
        t = ns.bottomup.BottomUpGenerateTarget(is_train=True, config=cfg, sigma=2.0, max_num=30)
        kps = bottomup_people(seed, m, 17, sizes)
        res = t.transform(dict(keypoints=[k.copy() for k in kps]))
        out[f"target_{ci}"] = res["target"]
        out[f"tag_ind_{ci}"] = res["tag_ind"]
    np.savez_compressed(os.path.join(golden_dir, "bottomup_encode_ref.npz"), **out)


def refine_inputs(seed, k=17, h=32, w=32, people=5):
    """heat / tag maps with `people` tagged blobs per joint and grouped keypoints [P,K,4] with
    some joints missing (val = 0), one plane without positive values, peaks on the border."""
    rng = np.random.RandomState(seed)
    heat = (rng.random_sample((k, h, w)) * 0.05).astype(np.float32)
    tagm = rng.uniform(-1, 1, (k, h, w, 1)).astype(np.float32)
    kps = np.zeros((people, k, 4), np.float32)
    bump = np.array([[.5, .8, .5], [.8, 1, .8], [.5, .8, .5]], np.float32)
    for p in range(people):
        ptag = 3.0 * p + rng.uniform(0, 0.5)
        for j in range(k):
            y, x = rng.randint(0, h), rng.randint(0, w)
            y0, y1, x0, x1 = max(y - 1, 0), min(y + 2, h), max(x - 1, 0), min(x + 2, w)
            heat[j, y0:y1, x0:x1] += rng.uniform(0.3, 0.9) * bump[y0 - y + 1:y1 - y + 1, x0 - x + 1:x1 - x + 1]
            tagm[j, max(y - 2, 0):y + 3, max(x - 2, 0):x + 3, 0] = ptag + rng.normal(0, 0.1)
            if rng.random_sample() < 0.6:
                kps[p, j] = (x, y, heat[j, y, x], tagm[j, y, x, 0])
        if not (kps[p, :, 2] > 0).any():
            kps[p, 0] = (3, 3, 0.5, ptag)
    heat[3] = -np.abs(heat[3])          # no positive value: nothing is filled in for joint 3
    heat[5] = np.round(heat[5] * 4) / 4  # plateaus: ties for the argmax and the +-0.25 shift
    return heat, tagm, kps


def refine_missing_golden(golden_dir):
    from oracle import refine_missing as rm

    f = rm.reference_function()
    out = {}
    for seed in range(4):
        heat, tagm, kps = refine_inputs(seed, people=[1, 3, 6, 11][seed])
        out[f"heat_{seed}"], out[f"tag_{seed}"], out[f"kps_{seed}"] = heat, tagm, kps
        out[f"refined_{seed}"] = np.stack([f(None, heat, tagm, kp.copy()) for kp in kps])
    np.savez_compressed(os.path.join(golden_dir, "refine_missing_ref.npz"), **out)


# (source h, w) of the eval-side preprocessing cases; max_image_size is [104, 64] so that the
# fixture stays small (the arithmetic does not depend on the size)
RESCALE_CASES = [(60, 80), (80, 60), (48, 101), (33, 37), (64, 128), (128, 208), (52, 52)]
RESCALE_MAX = [104, 64]


def rescale_image(seed, h, w):
    return np.random.RandomState(seed).randint(0, 256, (h, w, 3)).astype(np.uint8)


def rescale_pad_golden(ns, golden_dir):
    """BottomUpRescale -> BottomUpPad and BottomUpResize of the unmodified reference
    (bottomup_transform.py:144-209, :602-648, :212-302)."""
    import hashlib

    cfg = dict(image_size=[64, 64], max_image_size=RESCALE_MAX, heatmap_sizes=[[16, 16], [32, 32]],
               flip_pairs=[[1, 2]], pixel_std=200.0, tag_per_joint=True)
    rescale = ns.bottomup.BottomUpRescale(is_train=False, config=cfg)
    padder = ns.bottomup.BottomUpPad(is_train=False, config=cfg)
    resize = ns.bottomup.BottomUpResize(is_train=False, config=cfg, size=64, base_length=32)
    out = {}
    for ci, (h, w) in enumerate(RESCALE_CASES):
        img = rescale_image(100 + ci, h, w)
        r = rescale.transform(dict(image=img.copy()))
        p = padder.transform(dict(image=r["image"].copy()))
        out[f"rescaled_{ci}"] = r["image"]
        out[f"center_{ci}"], out[f"scale_{ci}"] = np.asarray(r["center"]), np.asarray(r["scale"])
        out[f"shape_{ci}"] = np.asarray(r["image_shape"])
        out[f"padded_{ci}"], out[f"mask_{ci}"] = p["image"], p["mask"]
        z = resize.transform(dict(image=img.copy()))
        out[f"resized_{ci}"], out[f"resized_mask_{ci}"] = z["image"], z["mask"]
        out[f"resized_center_{ci}"] = np.asarray(z["center"])
        out[f"resized_scale_{ci}"] = np.asarray(z["scale"])
        out[f"resized_shape_{ci}"] = np.asarray(z["image_shape"])
    # the shipped recipe's own sizes (max_image_size [832, 512]), kept as digests
    big = dict(cfg, max_image_size=[832, 512])
    rescale = ns.bottomup.BottomUpRescale(is_train=False, config=big)
    padder = ns.bottomup.BottomUpPad(is_train=False, config=big)
    digests = []
    for ci, (h, w) in enumerate([(480, 640), (640, 427), (375, 500)]):
        r = rescale.transform(dict(image=rescale_image(200 + ci, h, w)))
        p = padder.transform(dict(image=r["image"]))
        digests.append(hashlib.sha256(p["image"].tobytes() + p["mask"].tobytes()).hexdigest())
    out["big_digests"] = np.array(digests)
    np.savez_compressed(os.path.join(golden_dir, "bottomup_rescale_ref.npz"), **out)


def main(ns, golden_dir):
    bottomup_encode_golden(ns, golden_dir)
    refine_missing_golden(golden_dir)
    rescale_pad_golden(ns, golden_dir)

    import scipy.optimize

    # ---- scipy LSAP on tie-heavy matrices
    rng = np.random.RandomState(3)
    shapes, flat, rows, cols = [], [], [], []
    for it in range(300):
        nr, nc = rng.randint(1, 14), rng.randint(1, 16)
        if it % 3 == 0:
            c = rng.randint(0, 3, (nr, nc)).astype(np.float64)
        elif it % 3 == 1:
            c = np.round(rng.uniform(0, 5, (nr, nc)))
            if nc < nr:
                c = np.concatenate([c, np.zeros((nr, nr - nc)) + 1e10], axis=1)
        else:
            c = rng.uniform(0, 5, (nr, nc))
        r, cc = scipy.optimize.linear_sum_assignment(c)
        shapes.append(c.shape)
        flat.append(c.reshape(-1))
        rows.append(r)
        cols.append(cc)
    np.savez_compressed(
        os.path.join(golden_dir, "lsap_ref.npz"), shapes=np.array(shapes),
        cost=np.concatenate(flat), rows=np.concatenate(rows), cols=np.concatenate(cols),
        scipy_version=np.array(scipy.__version__))

    # ---- reference match_by_tag
    out = {}
    for mode, seed in (("people", 1), ("ties", 2), ("crowded", 3)):
        val, tag, ind = grouping_inputs(seed, 24, mode=mode)
        out[f"val_{mode}"], out[f"tag_{mode}"], out[f"ind_{mode}"] = val, tag, ind
        for rounded in (True, False):
            res = [ns.match.match_by_tag(v, t, x, synth.COCO_JOINT_ORDER, vis_thr=0.1, tag_thr=1.0,
                                         ignore_too_much=False, use_rounded_norm=rounded)
                   for v, t, x in zip(val, tag, ind)]
            counts = np.array([0 if r.ndim == 1 else r.shape[0] for r in res])
            body = [r.reshape(-1) for r in res if r.ndim == 3]
            name = f"{mode}_{'rounded' if rounded else 'exact'}"
            out[f"counts_{name}"] = counts
            out[f"ans_{name}"] = np.concatenate(body) if body else np.zeros(0, np.float32)
    np.savez_compressed(os.path.join(golden_dir, "match_ref.npz"), **out)

    # ---- restated bottom-up decode, frozen
    from oracle import bottomup_decode as bd

    d = synth.bottomup_outputs(2, 17, 32, 32, mask_hw=(128, 128), seed=4, max_people=4)
    val_k, tag_k, ind_k, raw, tagging = bd.decode([d["out0"], d["out1"]], d["mask"], use_nms=True,
                                                  nms_kernel=3, max_num=30)
    np.savez_compressed(os.path.join(golden_dir, "bottomup_decode_restated.npz"),
                        val_k=val_k, tag_k=tag_k, ind_k=ind_k,
                        raw_sum=raw.sum(axis=(2, 3)), tag_sum=tagging.sum(axis=(2, 3, 4)))
