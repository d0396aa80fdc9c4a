"""Drop-in top-down transform classes (registry module ``transform``).

They keep the reference's calling convention -- constructed as
``cls(is_train=..., config=dataset_setting, **yaml_kwargs)``, invoked per sample
as ``t(*columns) -> tuple[np.ndarray, ...]`` with the columns of
``COLUMN_MAP`` (mindpose/data/transform/transform.py:66-79) -- and add batched
device entry points (``*_batch``), which is how the codec reaches HBM speed: a
per-sample launch cannot.  Both routes run the same CUDA kernels.
"""
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import codec
from .column_names import COLUMN_MAP
from .register import register


def _dev() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "mindpose_b200 transforms run on a CUDA device; none is visible "
            "(there is no CPU fallback)"
        )
    return torch.device("cuda", torch.cuda.current_device())


def _three_point_geometry(center, scale, rot, out_wh, pixel_std):
    """The two point triples ``get_affine_transform`` hands to cv2.getAffineTransform
    (utils.py:73-96), float32 [3,2] each, evaluated with numpy in the dtypes the caller passed
    (what the reference does: the scalars are float64 or float32 as they arrive, the points are
    stored as float32)."""
    w, h = out_wh
    half_w = (scale * pixel_std)[0] * -0.5
    angle = np.pi * rot / 180
    sn, cs = np.sin(angle), np.cos(angle)
    direction = [0.0 * cs - half_w * sn, 0.0 * sn + half_w * cs]
    src = np.zeros((3, 2), dtype=np.float32)
    src[0] = center
    src[1] = center + direction
    dst = np.zeros((3, 2), dtype=np.float32)
    dst[0] = [w * 0.5, h * 0.5]
    dst[1] = np.array([w * 0.5, h * 0.5]) + np.array([0.0, w * -0.5])
    for pts in (src, dst):  # third point: (a - b) turned by 90 degrees about b
        d = pts[0] - pts[1]
        pts[2] = pts[1] + np.array([-d[1], d[0]], dtype=np.float32)
    return src, dst


class Transform:
    """Column tuple <-> state dict adapter (transform.py:6-79)."""

    def __init__(self, is_train: bool = True, config: Optional[Dict[str, Any]] = None) -> None:
        self.is_train = is_train
        self.config = config if config else dict()
        self._transform_cfg = self.load_transform_cfg()
        self._required_field = self.setup_required_field()

    def load_transform_cfg(self) -> Dict[str, Any]:
        raise NotImplementedError("Child class must implement this method.")

    def transform(self, state: Dict[str, Any]) -> Dict[str, Any]:
        raise NotImplementedError("Child class must implement this method.")

    def setup_required_field(self) -> List[str]:
        raise NotImplementedError("Child class must implement this method.")

    def __call__(self, *args: Any) -> Tuple[np.ndarray, ...]:
        state = dict(zip(self._required_field, args))
        state.update(self.transform(state))
        return tuple(np.asarray(state[name]) for name in self._required_field)


class TopDownTransform(Transform):
    """Shared config parsing of the top-down transforms
    (topdown_transform.py:32-94)."""

    def setup_required_field(self) -> List[str]:
        return COLUMN_MAP["topdown"]["train" if self.is_train else "val"]

    def load_transform_cfg(self) -> Dict[str, Any]:
        cfg = dict()
        cfg["image_size"] = np.array(self.config["image_size"])
        cfg["heatmap_size"] = np.array(self.config["heatmap_size"])
        assert len(cfg["image_size"]) == 2
        assert len(cfg["heatmap_size"]) == 2
        pairs = np.array(self.config["flip_pairs"])
        if pairs.ndim == 2:
            cfg["flip_index"] = np.insert(pairs[:, ::-1].flatten(), 0, 0)
        else:
            cfg["flip_index"] = pairs
        cfg["flip_pairs"] = pairs
        cfg["upper_body_ids"] = np.array(self.config["upper_body_ids"])
        cfg["pixel_std"] = float(self.config["pixel_std"])
        cfg["scale_padding"] = float(self.config["scale_padding"])
        jw = self.config.get("joint_weights")
        cfg["joint_weights"] = None if jw is None else np.array(jw)
        return cfg


@register("transform", extra_name="topdown_box_to_center_scale")
class TopDownBoxToCenterScale(TopDownTransform):
    """Box (x, y, w, h) -> center / scale (topdown_transform.py:97-154).

    Required keys: boxes.  Returned keys: center, scale.
    The training-time random centre shift is host-side augmentation outside the
    codec; it is applied to the box before the kernel exactly where the
    reference applies it to the centre (:140-141).
    """

    def transform(self, state: Dict[str, Any]) -> Dict[str, Any]:
        boxes = np.asarray(state["boxes"], dtype=np.float32).reshape(1, 4)
        center, scale = self.box_to_center_scale_batch(torch.from_numpy(boxes).to(_dev()))
        center = center[0].cpu().numpy()
        if self.is_train and np.random.rand() < 0.3:
            w, h = boxes[0, 2], boxes[0, 3]
            center += np.random.uniform(-0.2, 0.2, size=2) * [w, h]
        return dict(center=center, scale=scale[0].cpu().numpy())

    def box_to_center_scale_batch(self, boxes: torch.Tensor):
        return codec.box_to_center_scale(
            boxes,
            self._transform_cfg["image_size"],
            pixel_std=self._transform_cfg["pixel_std"],
            scale_padding=self._transform_cfg["scale_padding"],
        )


@register("transform", extra_name="topdown_affine")
class TopDownAffine(TopDownTransform):
    """Affine crop warp of one instance (topdown_transform.py:157-261).

    Required keys: image, center, scale, rotation, keypoints (optional).
    Returned keys: image, keypoints (optional).
    """

    def __init__(self, is_train: bool = True, config: Optional[Dict[str, Any]] = None,
                 use_udp: bool = False) -> None:
        super().__init__(is_train=is_train, config=config)
        self.use_udp = use_udp

    def _host_geometry(self, center, scale, rot):
        """The rotation-dependent scalars of ONE sample whose geometry lives on the host.

        With a rotation the reference's matrix depends on numpy's own sin / cos and on the
        dtype the rotation arrives in (a float32 scalar from the dataset pipeline; numpy then
        rounds the angle, and under NumPy >= 2 the rotated direction, in float32).  Those few
        scalars are evaluated here with numpy -- in whatever dtypes the caller passed, like the
        reference does (utils.py:73-96 / :158-190).
        -> ("points", src f32 [3,2], dst f32 [3,2]) for cv2.getAffineTransform, or
           ("matrix", m f32 [2,3]) for UDP."""
        cfg = self._transform_cfg
        w, h = cfg["image_size"]
        pixel_std = cfg["pixel_std"]
        if self.use_udp:
            theta = np.deg2rad(rot)
            size_in, size_dst = center * 2.0, np.asarray(cfg["image_size"]) - 1.0
            size_tgt = scale * pixel_std
            kx, ky = size_dst[0] / size_tgt[0], size_dst[1] / size_tgt[1]
            cs, sn = np.cos(theta), np.sin(theta)
            m = np.zeros((2, 3), dtype=np.float32)
            m[0, 0], m[0, 1] = cs * kx, -sn * kx
            m[0, 2] = kx * (-0.5 * size_in[0] * cs + 0.5 * size_in[1] * sn + 0.5 * size_tgt[0])
            m[1, 0], m[1, 1] = sn * ky, cs * ky
            m[1, 2] = ky * (-0.5 * size_in[0] * sn - 0.5 * size_in[1] * cs + 0.5 * size_tgt[1])
            return "matrix", m
        return ("points",) + _three_point_geometry(center, scale, rot, (w, h), pixel_std)

    def _host_matrix(self, center, scale, rot) -> Optional[torch.Tensor]:
        """Forward matrix f64 [1,2,3] on the device for a ROTATED sample (None at rot == 0,
        where the device's own matrix is bit-exact): host scalars from `_host_geometry`, the
        3-point solve (and later the warp and the joints) on the device."""
        if not np.any(np.asarray(rot) != 0):
            return None
        dev = _dev()
        geo = self._host_geometry(center, scale, rot)
        if geo[0] == "matrix":
            return torch.from_numpy(geo[1].astype(np.float64)[None]).to(dev)
        fwd, _ = codec.affine_from_points(torch.from_numpy(geo[1][None]).to(dev),
                                          torch.from_numpy(geo[2][None]).to(dev))
        return fwd

    def transform(self, state: Dict[str, Any]) -> Dict[str, Any]:
        dev = _dev()
        image = np.ascontiguousarray(state["image"])
        if image.dtype != np.uint8 or image.ndim != 3:
            raise ValueError("`image` must be a uint8 HWC array")
        center = torch.from_numpy(np.asarray(state["center"], np.float32).reshape(1, 2)).to(dev)
        scale = torch.from_numpy(np.asarray(state["scale"], np.float32).reshape(1, 2)).to(dev)
        rot = torch.from_numpy(np.asarray(state["rotation"], np.float32).reshape(1)).to(dev)
        kps = None
        if "keypoints" in state:
            kps = torch.from_numpy(
                np.ascontiguousarray(state["keypoints"], dtype=np.float32)[None]).to(dev)
        img_t = torch.from_numpy(image[None]).to(dev)
        fwd = self._host_matrix(state["center"], state["scale"], state["rotation"])
        crop, kps = self.affine_batch(img_t, center, scale, rot, kps, matrices=fwd)
        out = dict(image=crop[0].cpu().numpy())
        if kps is not None:
            # the reference mutates state["keypoints"] in place
            state["keypoints"][...] = kps[0].cpu().numpy().astype(state["keypoints"].dtype)
            out["keypoints"] = state["keypoints"]
        return out

    def affine_batch(self, images: torch.Tensor, center, scale, rot=None, keypoints=None,
                     normalize_mean=None, normalize_std=None, matrices=None):
        """images u8 [N,Hs,Ws,C] (one per crop) -> (crops u8 [N,h,w,C], keypoints).

        ``matrices`` (f64 [N,2,3], CUDA): forward matrices to use instead of the ones the
        device derives from (center, scale, rot) -- e.g. made by ``codec.affine_from_points``.

        With ``normalize_mean`` / ``normalize_std`` (the ``create_pipeline`` arguments, in
        [0, 1] units: data_factory.py:78-79) the pipeline's Normalize + HWC2CHW step is fused
        into the warp and the crops come back as float32 [N,3,h,w]."""
        cfg = self._transform_cfg
        if matrices is not None:
            fwd = matrices.to(torch.float64).reshape(-1, 2, 3).contiguous()
            inv = codec.invert_affine(fwd)
        else:
            fwd, inv = codec.affine_matrices(center, scale, rot, cfg["image_size"],
                                             pixel_std=cfg["pixel_std"], use_udp=self.use_udp)
        if normalize_mean is not None:
            n, hs, ws, c = images.shape
            off = torch.arange(n, device=images.device, dtype=torch.int64) * (hs * ws * c)
            hw = torch.tensor([hs, ws], device=images.device, dtype=torch.int32).repeat(n, 1)
            # np.array(mean) * 255.0 in float64, handed to Normalize as float32
            mean = (np.array(normalize_mean) * 255.0).tolist()
            std = (np.array(normalize_std) * 255.0).tolist()
            crops = codec.warp_affine_normalized(images, off, hw, inv, cfg["image_size"], mean, std)
        else:
            crops = codec.warp_affine_uniform(images, inv, cfg["image_size"])
        if keypoints is not None:
            keypoints = codec.affine_joints(keypoints.contiguous(), fwd, use_udp=self.use_udp)
        return crops, keypoints


@register("transform", extra_name="topdown_generate_target")
class TopDownGenerateTarget(TopDownTransform):
    """Keypoints -> Gaussian heatmap targets (topdown_transform.py:264-430).

    Required keys: keypoints.  Returned keys: target, target_weight.
    """

    def __init__(self, is_train: bool = True, config: Optional[Dict[str, Any]] = None,
                 sigma: float = 2.0, use_different_joint_weights: bool = False,
                 use_udp: bool = False) -> None:
        super().__init__(is_train=is_train, config=config)
        self.sigma = sigma
        self.use_different_joint_weights = use_different_joint_weights
        self.use_udp = use_udp
        if self.use_different_joint_weights and self._transform_cfg["joint_weights"] is None:
            raise ValueError(
                "`joint_weights` must be provided if `use_different_joint_weights` is True."
            )

    def transform(self, state: Dict[str, Any]) -> Dict[str, Any]:
        kps = torch.from_numpy(
            np.ascontiguousarray(state["keypoints"], dtype=np.float32)[None]).to(_dev())
        target, weight = self.encode_batch(kps)
        return dict(target=target[0].cpu().numpy(), target_weight=weight[0].cpu().numpy())

    def encode_batch(self, keypoints: torch.Tensor, out: Optional[torch.Tensor] = None):
        """keypoints f32 [N,K,3] -> (target [N,K,H,W], target_weight [N,K])."""
        cfg = self._transform_cfg
        jw = cfg["joint_weights"] if self.use_different_joint_weights else None
        return codec.topdown_encode(keypoints, cfg["image_size"], cfg["heatmap_size"],
                                    sigma=self.sigma, use_udp=self.use_udp, joint_weights=jw,
                                    out=out)


class BottomUpTransform(Transform):
    """Shared config parsing of the bottom-up transforms (bottomup_transform.py:24-85)."""

    def setup_required_field(self) -> List[str]:
        return COLUMN_MAP["bottomup"]["train" if self.is_train else "val"]

    def load_transform_cfg(self) -> Dict[str, Any]:
        cfg = dict()
        cfg["image_size"] = np.array(self.config["image_size"])
        cfg["max_image_size"] = np.array(self.config["max_image_size"])
        cfg["heatmap_sizes"] = np.array(self.config["heatmap_sizes"])
        assert len(cfg["image_size"]) == 2
        for x in cfg["heatmap_sizes"]:
            assert len(x) == 2
        pairs = np.array(self.config["flip_pairs"])
        if pairs.ndim == 2:
            cfg["flip_index"] = np.insert(pairs[:, ::-1].flatten(), 0, 0)
        else:
            cfg["flip_index"] = pairs
        cfg["flip_pairs"] = pairs
        cfg["pixel_std"] = float(self.config["pixel_std"])
        cfg["tag_per_joint"] = self.config["tag_per_joint"]
        return cfg


@register("transform", extra_name="bottomup_generate_target")
class BottomUpGenerateTarget(BottomUpTransform):
    """Keypoints -> multi-resolution Gaussian heatmaps + tag indices
    (bottomup_transform.py:463-598).

    Required keys: keypoints (one [M,K,3] array per scale).  Returned keys: target, tag_ind.
    """

    def __init__(self, is_train: bool = True, config: Optional[Dict[str, Any]] = None,
                 sigma: float = 2.0, max_num: int = 30) -> None:
        super().__init__(is_train=is_train, config=config)
        self.sigma = sigma
        self.max_num = max_num

    def transform(self, state: Dict[str, Any]) -> Dict[str, Any]:
        kps = [np.asarray(k, dtype=np.float32) for k in state["keypoints"]]
        m = kps[0].shape[0]
        if m > self.max_num:
            raise ValueError(
                f"Number of keypoints in one image `{m}` exeeds the maximum num: `{self.max_num}`")
        batch = torch.from_numpy(np.ascontiguousarray(np.stack(kps)[None])).to(_dev())
        target, tag_ind = self.encode_batch(batch)
        return dict(target=target[0].cpu().numpy(), tag_ind=tag_ind[0].cpu().numpy())

    def encode_batch(self, keypoints: torch.Tensor):
        """keypoints f32 [N,S,M,K,3] -> (target [N,S,K,Hmax,Wmax], tag_ind [N,S,max_num,K,2])."""
        from . import bottomup

        cfg = self._transform_cfg
        return bottomup.encode_targets(keypoints, cfg["heatmap_sizes"], sigma=self.sigma,
                                       max_num=self.max_num, tag_per_joint=cfg["tag_per_joint"])


def _rescale_size(image_wh, max_wh) -> Tuple[int, int]:
    """Target (w, h) of ``BottomUpRescale`` (bottomup_transform.py:152-168): the image fills the
    ``max_image_size`` box (turned by 90 degrees for a portrait image) along one side; float64
    ratios and Python's round (half to even), as the reference evaluates them."""
    w, h = int(image_wh[0]), int(image_wh[1])
    box_w, box_h = int(max_wh[0]), int(max_wh[1])
    if w < h:
        box_w, box_h = box_h, box_w
    if w / h > box_w / box_h:
        return box_w, int(round(h * box_w / w))
    return int(round(w * box_h / h)), box_h


@register("transform", extra_name="bottomup_rescale")
class BottomUpRescale(BottomUpTransform):
    """Rescale the image into the ``max_image_size`` box, aspect ratio kept
    (bottomup_transform.py:144-209; the first validation transform of the shipped recipe).

    Required keys: image.  Returned keys: image, center, scale, image_shape.
    ``rescale_pad_batch`` runs this transform and ``BottomUpPad`` for a batch in one kernel.
    """

    def transform(self, state: Dict[str, Any]) -> Dict[str, Any]:
        image = np.ascontiguousarray(state["image"])
        if image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] != 3:
            raise ValueError("`image` must be a uint8 HWC array of 3 channels")
        height, width = image.shape[:2]
        tw, th = _rescale_size((width, height), self._transform_cfg["max_image_size"])
        canvas, _, meta = self.rescale_pad_batch([image], canvas_wh=(tw, th), with_mask=False)
        return dict(image=canvas[0].cpu().numpy(), center=meta["center"][0],
                    scale=meta["scale"][0], image_shape=(tw, th))

    def rescale_pad_batch(self, images, canvas_wh=None, with_mask: bool = True,
                          normalize_mean=None, normalize_std=None):
        """images: a list of uint8 HWC arrays / tensors (any sizes), or one u8 tensor
        [N, H, W, 3] -> (canvas u8 [N, CH, CW, 3] CUDA, mask u8 [N, CH, CW] or None, meta) with
        meta = dict(center int64 [N,2], scale f64 [N,2], image_shape int64 [N,2]) as numpy,
        exactly what BottomUpRescale followed by BottomUpPad return per image.

        ``canvas_wh`` defaults to ``max_image_size`` (landscape box); portrait images need the
        turned box, so a mixed batch needs a canvas that holds both (the reference pads each
        image to its own orientation; a batch tensor has one shape).

        With ``normalize_mean`` / ``normalize_std`` (the ``create_pipeline`` arguments, in [0, 1]
        units: data_factory.py:78-79) the pipeline's Normalize + HWC2CHW step is fused in and
        the canvas comes back as float32 [N, 3, CH, CW] (padding = (0 - mean) / std)."""
        cfg = self._transform_cfg
        dev = _dev()
        if isinstance(images, torch.Tensor) and images.dim() == 4:
            images = list(images)
        sizes = [(int(im.shape[1]), int(im.shape[0])) for im in images]   # (w, h)
        for im in images:
            if im.ndim != 3 or im.shape[2] != 3:
                raise ValueError("every image must be HWC with 3 channels")
        targets = [_rescale_size(wh, cfg["max_image_size"]) for wh in sizes]
        if canvas_wh is None:
            canvas_wh = tuple(int(v) for v in cfg["max_image_size"])
        cw, ch = int(canvas_wh[0]), int(canvas_wh[1])
        for (tw, th) in targets:
            if tw > cw or th > ch:   # the reference: assert target_width >= width ...
                raise ValueError(f"a {tw}x{th} rescaled image does not fit the {cw}x{ch} canvas")
        # one allocation, every image at a 16-byte aligned offset
        offs, total = [], 0
        for (w, h) in sizes:
            offs.append(total)
            total += (w * h * 3 + 15) // 16 * 16
        blob = torch.empty(max(total, 16), dtype=torch.uint8, device=dev)
        for im, off, (w, h) in zip(images, offs, sizes):
            t = im if isinstance(im, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(im))
            if t.dtype != torch.uint8:
                raise ValueError("images must be uint8")
            blob[off:off + w * h * 3].copy_(t.reshape(-1), non_blocking=True)
        src_hw = torch.tensor([[h, w] for (w, h) in sizes], dtype=torch.int32).reshape(-1, 2)
        dst_wh = torch.tensor(targets, dtype=torch.int32).reshape(-1, 2)
        mean = std = None
        if normalize_mean is not None:
            # np.array(mean) * 255.0 in float64, handed to Normalize as float32
            mean = (np.array(normalize_mean) * 255.0).tolist()
            std = (np.array(normalize_std) * 255.0).tolist()
        canvas, mask = codec.rescale_pad(blob, torch.tensor(offs, dtype=torch.int64), src_hw,
                                         dst_wh, (cw, ch), with_mask=with_mask, mean=mean, std=std)
        pixel_std = cfg["pixel_std"]
        meta = dict(center=np.array([[round(w / 2), round(h / 2)] for (w, h) in sizes]),
                    scale=np.array([[w / pixel_std, h / pixel_std] for (w, h) in sizes]),
                    image_shape=np.array(targets))
        return canvas, mask, meta


@register("transform", extra_name="bottomup_pad")
class BottomUpPad(BottomUpTransform):
    """Pad the image with zeros to ``max_image_size`` (turned for a portrait image) and make the
    validity mask (bottomup_transform.py:602-648).  Required keys: image.  Returned keys:
    image, mask.  A copy, not arithmetic: it stays on the host for a single sample;
    ``BottomUpRescale.rescale_pad_batch`` writes the padded canvas and the mask on the device
    in the same pass as the rescale."""

    def transform(self, state: Dict[str, Any]) -> Dict[str, Any]:
        image = state["image"]
        height, width = image.shape[:2]
        target_w, target_h = (int(v) for v in self._transform_cfg["max_image_size"])
        if width < height:
            target_w, target_h = target_h, target_w
        if target_w < width or target_h < height:
            raise AssertionError("the image is larger than max_image_size")
        out = np.zeros((target_h, target_w) + tuple(image.shape[2:]), dtype=image.dtype)
        out[:height, :width] = image
        mask = np.zeros((target_h, target_w), dtype=np.uint8)
        mask[:height, :width] = 1
        return dict(image=out, mask=mask)


@register("transform", extra_name="bottomup_resize")
class BottomUpResize(BottomUpTransform):
    """Short side to ``size`` (rounded up to ``base_length``), long side rounded up to
    ``base_length``, through the affine warp (bottomup_transform.py:212-302: get_affine_transform
    with rotation 0 and cv2.warpAffine) -- the crop-warp kernel of the top-down path.

    Required keys: image.  Returned keys: image, mask, center, scale, image_shape.
    """

    def __init__(self, is_train: bool = True, config: Optional[Dict[str, Any]] = None,
                 size: int = 512, base_length: int = 64) -> None:
        super().__init__(is_train=is_train, config=config)
        self.size = size
        self.base_length = base_length

    def _get_new_size(self, image_wh, pixel_std: float = 200.0):
        w, h = image_wh
        unit = self.base_length

        def up(x):
            return int(np.ceil(x / unit)) * unit

        short = up(self.size)
        if w < h:
            tw, th = short, up(short / w * h)
            scale = np.array([w / pixel_std, th / tw * w / pixel_std])
        else:
            th, tw = short, up(short / h * w)
            scale = np.array([tw / th * h / pixel_std, h / pixel_std])
        return (tw, th), np.array([round(w / 2), round(h / 2)]), scale

    def transform(self, state: Dict[str, Any]) -> Dict[str, Any]:
        dev = _dev()
        image = np.ascontiguousarray(state["image"])
        if image.dtype != np.uint8 or image.ndim != 3:
            raise ValueError("`image` must be a uint8 HWC array")
        height, width = image.shape[:2]
        target, center, scale = self._get_new_size((width, height),
                                                   self._transform_cfg["pixel_std"])
        # the reference calls get_affine_transform with its default pixel_std (200), whatever
        # the config says (bottomup_transform.py:291)
        src, dst = _three_point_geometry(center, scale, 0.0, target, 200.0)
        _, inv = codec.affine_from_points(torch.from_numpy(src[None]).to(dev),
                                          torch.from_numpy(dst[None]).to(dev))
        out = codec.warp_affine_uniform(torch.from_numpy(image[None]).to(dev), inv, target)
        return dict(image=out[0].cpu().numpy(),
                    mask=np.ones((target[1], target[0]), dtype=np.uint8),
                    center=center, scale=scale, image_shape=target)
