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
        return "points", src, dst

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
