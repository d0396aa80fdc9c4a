"""Drop-in inference engines (registry module ``inferencer``): only the codec
half -- flip-test averaging + decode (top-down), tag grouping + scoring +
back-projection (bottom-up).  The network is any callable returning heatmaps on
the CUDA device; the backbone stays outside the graft.
"""
from typing import Any, Callable, Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch

from .decoders import BottomUpHeatMapAEDecoder, TopDownHeatMapDecoder
from .register import register


class Inferencer:
    """Abstract engine (mindpose/engine/inferencer/inferencer.py)."""

    def __init__(self, net: Callable, config: Optional[Dict[str, Any]] = None) -> None:
        self.net = net
        self.config = config if config else dict()
        self._inference_cfg = self.load_inference_cfg()

    def load_inference_cfg(self) -> Dict[str, Any]:
        raise NotImplementedError("Child class must implement this method.")

    def infer(self, dataset) -> List[Dict[str, Any]]:
        raise NotImplementedError("Child class must implement this method.")

    def __call__(self, dataset) -> List[Dict[str, Any]]:
        return self.infer(dataset)


def _flip_index(flip_pairs) -> np.ndarray:
    return np.insert(np.array(flip_pairs)[:, ::-1].flatten(), 0, 0)


class _MultiRunNet:
    """Horizontal-flip test-time augmentation, top-down
    (mindpose/engine/inferencer/topdown_inferencer.py:146-187).

    ``net(image)`` must return the heatmap [N,K,H,W].  The two network outputs go
    straight into the fused kernel: channel permutation, x reversal, optional
    1-px shift, averaging and decoding happen in one read of each stack.
    """

    def __init__(self, net: Callable, decoder: TopDownHeatMapDecoder, flip_index,
                 shift_heatmap: bool = False) -> None:
        self.net = net
        self.decoder = decoder
        self.shift_heatmap = shift_heatmap
        self.flip_index = np.asarray(flip_index)

    def set_train(self, mode: bool = False) -> "_MultiRunNet":
        return self

    def construct(self, image: torch.Tensor, center, scale, score) -> Tuple[torch.Tensor, torch.Tensor]:
        heatmap = self.net(image)
        flipped_heatmap = self.net(torch.flip(image, dims=[3]))
        return self.decoder.decode_flip_pair(heatmap, flipped_heatmap, self.flip_index, center,
                                             scale, score, shift_heatmap=self.shift_heatmap)

    __call__ = construct


@register("inferencer", extra_name="topdown_heatmap")
class TopDownHeatMapInferencer(Inferencer):
    """Top-down heatmap inference over a dataset
    (mindpose/engine/inferencer/topdown_inferencer.py:16-144).

    Args:
        net: callable ``image -> heatmap [N,K,H,W]`` on the CUDA device
        config: needs ``has_heatmap_output``, ``hflip_tta``, ``shift_heatmap``, ``flip_pairs``
        progress_bar: kept for signature parity
        decoder: TopDownHeatMapDecoder

    ``dataset`` is an iterable of dicts with keys image, center, scale,
    bbox_scores, image_file, bbox_ids.  Returns one record per crop with keys
    pred, box, image_path, bbox_id.
    """

    def __init__(self, net: Callable, config: Optional[Dict[str, Any]] = None,
                 progress_bar: bool = False,
                 decoder: Optional[TopDownHeatMapDecoder] = None) -> None:
        super().__init__(net, config=config)
        self.progress_bar = progress_bar
        self.decoder = decoder
        if self.decoder is None and self._inference_cfg["hflip_tta"]:
            raise ValueError("Decoder must be provided for flip TTA")
        if self._inference_cfg["hflip_tta"] and not self._inference_cfg["has_heatmap_output"]:
            raise ValueError("flip TTA need heatmap output.")
        if self.decoder is None:
            raise ValueError("Decoder must be provided")
        self._multi_run_net = None
        if self._inference_cfg["hflip_tta"]:
            self._multi_run_net = _MultiRunNet(
                self.net, self.decoder, self._inference_cfg["flip_index"],
                shift_heatmap=self._inference_cfg["shift_heatmap"])

    def load_inference_cfg(self) -> Dict[str, Any]:
        return dict(
            has_heatmap_output=self.config["has_heatmap_output"],
            hflip_tta=self.config["hflip_tta"],
            shift_heatmap=self.config["shift_heatmap"],
            flip_index=_flip_index(self.config["flip_pairs"]),
        )

    def infer(self, dataset: Iterable[Dict[str, Any]]) -> List[Dict[str, Any]]:
        outputs = list()
        for data in dataset:
            if self._multi_run_net is not None:
                preds, boxes = self._multi_run_net(
                    data["image"], data["center"], data["scale"], data["bbox_scores"])
            else:
                preds, boxes = self.decoder(
                    self.net(data["image"]), data["center"], data["scale"], data["bbox_scores"])
            preds = preds.cpu().numpy()
            boxes = boxes.cpu().numpy()
            for pred, box, path, bbox_id in zip(preds, boxes, data["image_file"], data["bbox_ids"]):
                outputs.append(dict(pred=pred.tolist(), box=box.tolist(),
                                    image_path=np.asarray(path).tolist(),
                                    bbox_id=np.asarray(bbox_id).tolist()))
        return outputs


@register("inferencer", extra_name="bottomup_heatmap_ae")
class BottomUpHeatMapAEInferencer(Inferencer):
    """Bottom-up inference: decode, group by tag, score, back-project
    (mindpose/engine/inferencer/bottomup_inferencer.py:19-187).

    ``net(image) -> [out0, out1]``.  ``dataset`` yields dicts with keys image,
    mask, center, scale, image_shape, image_file.  Records: pred, score, image_path.
    """

    def __init__(self, net: Callable, config: Optional[Dict[str, Any]] = None,
                 progress_bar: bool = False,
                 decoder: Optional[BottomUpHeatMapAEDecoder] = None) -> None:
        super().__init__(net, config=config)
        self.progress_bar = progress_bar
        self.decoder = decoder
        if self._inference_cfg["hflip_tta"]:
            # the reference's bottom-up flip test multiplies Python lists
            # (bottomup_inferencer.py:274-281) and cannot run; not reproduced
            raise ValueError("bottom-up flip TTA is not supported")
        if self.decoder is None:
            raise ValueError("Decoder must be provided")

    def load_inference_cfg(self) -> Dict[str, Any]:
        c = self.config
        return dict(
            has_heatmap_output=c["has_heatmap_output"],
            hflip_tta=c["hflip_tta"],
            joint_order=c["joint_order"],
            vis_thr=float(c["vis_thr"]),
            ignore_too_much=c["ignore_too_much"],
            use_rounded_norm=c["use_rounded_norm"],
            tag_thr=float(c["tag_thr"]),
            pixel_std=float(c["pixel_std"]),
            downsample_scale=c["downsample_scale"],
            refine_missing_joint=c["refine_missing_joint"],
            flip_index=_flip_index(c["flip_pairs"]),
        )

    def infer(self, dataset: Iterable[Dict[str, Any]]) -> List[Dict[str, Any]]:
        from . import bottomup

        outputs = list()
        cfg = self._inference_cfg
        for data in dataset:
            val_k, tag_k, ind_k, raw, tagging = self.decoder(self.net(data["image"]), data["mask"])
            gkw = dict(joint_order=cfg["joint_order"], vis_thr=cfg["vis_thr"],
                       tag_thr=cfg["tag_thr"], ignore_too_much=cfg["ignore_too_much"],
                       use_rounded_norm=cfg["use_rounded_norm"])
            ans, num, scores = bottomup.group_by_tag(val_k, tag_k, ind_k, **gkw)
            if bool((num < 0).any()):
                # more than the default 128 people in some image (the reference is unbounded,
                # match.py:63-113): once more with room for every detection as its own group
                # (or what shared memory holds, when max_num > 32)
                room = bottomup.max_group_capacity(val_k.shape[1], val_k.shape[2])
                ans, num, scores = bottomup.group_by_tag(val_k, tag_k, ind_k, max_groups=room,
                                                         **gkw)
            if cfg["refine_missing_joint"]:  # after the scores, as the reference (:153-166)
                if raw is None or tagging is None:
                    raise ValueError("refine_missing_joint needs the decoder's heatmap outputs "
                                     "(decoder.return_maps = True)")
                bottomup.refine_missing(raw, tagging, ans, num)
            image_shape = torch.as_tensor(np.asarray(data["image_shape"]), dtype=torch.float64)
            bottomup.transform_keypoints(
                ans, num, data["center"], data["scale"],
                image_shape / cfg["downsample_scale"], pixel_std=cfg["pixel_std"])
            ans_h, num_h, scores_h = ans.cpu().numpy(), num.cpu().numpy(), scores.cpu().numpy()
            for i, path in enumerate(data["image_file"]):
                p = int(num_h[i])
                if p < 0:   # cannot happen with K * M groups of room
                    raise RuntimeError("more people in one image than group_by_tag can hold")
                outputs.append(dict(pred=ans_h[i, :p], score=scores_h[i, :p].tolist(),
                                    image_path=path))
        return outputs
