"""Drop-in decoder classes (registry module ``decoder``).

Same names, constructor keywords, call signatures, return arity and error types
as the reference decoders; the bodies are single calls into the fused CUDA
kernels.  "Tensor" here is a ``torch.Tensor`` on a CUDA device.
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import codec
from .register import register


class Decoder:
    """Abstract decoder (mindpose/models/decoders/decoder.py)."""

    def construct(self, *args, **kwargs):
        raise NotImplementedError("Child class must implement this method.")

    def __call__(self, *args, **kwargs):
        return self.construct(*args, **kwargs)

    def set_train(self, mode: bool = False) -> "Decoder":
        # decoders are stateless; kept so call sites written for nn.Cell work
        return self


@register("decoder", extra_name="topdown_heatmap")
class TopDownHeatMapDecoder(Decoder):
    """Decode top-down heatmaps into image coordinates and boxes.

    Mirrors ``TopDownHeatMapDecoder``
    (mindpose/models/decoders/top_down_decoder.py:13-215).

    Args:
        pixel_std: scaling factor used in decoding. Default: 200.
        to_original: map the coordinates back to the raw image. Default: True
        shift_coordinate: +-0.25 px shift toward the higher neighbour. Default: False
        use_udp: UDP back-projection (divide by size - 1). Default: False
        dark_udp_refine: DARK / UDP Taylor refinement; exclusive with
            ``shift_coordinate``. Default: False
        kernel_size: blur kernel of the refinement (11 for sigma 2, 17 for sigma 3).

    Inputs: heatmap [N,K,H,W], center [N,2], scale [N,2], score [N]
    Outputs: all_preds [N,K,3] (x, y, maxval), all_boxes [N,6]
        (center_x, center_y, scale_x, scale_y, area, score)
    """

    def __init__(
        self,
        pixel_std: float = 200.0,
        to_original: bool = True,
        shift_coordinate: bool = False,
        use_udp: bool = False,
        dark_udp_refine: bool = False,
        kernel_size: int = 11,
    ) -> None:
        self.pixel_std = pixel_std
        self.to_original = to_original
        self.shift_coordinate = shift_coordinate
        self.use_udp = use_udp
        self.dark_udp_refine = dark_udp_refine
        self.kernel_size = kernel_size

        if self.dark_udp_refine and self.shift_coordinate:
            raise ValueError(
                "`udp_refine` and `shift_coordinate` cannot be `true` in the same time."
            )
        self.gaussian_kernel = (
            self._create_gaussian_kernel(kernel_size) if dark_udp_refine else None
        )

    @staticmethod
    def _create_gaussian_kernel(kernel_size: int) -> np.ndarray:
        """float32 [1,1,ks,ks], sum 1 (top_down_decoder.py:207-215)."""
        half = (kernel_size - 1) // 2
        sigma = 0.3 * ((kernel_size - 1) * 0.5 - 1) + 0.8
        t = np.arange(-half, half + 1)
        k = np.exp(-(t[None, :] ** 2 + t[:, None] ** 2) / (2 * sigma**2))
        return (k / k.sum()).astype(np.float32)[None, None]

    def _params(self, k: int, h: int, w: int, flip_index=None, shift_heatmap=False):
        return codec.make_decode_params(
            k, h, w,
            pixel_std=self.pixel_std,
            to_original=self.to_original,
            shift_coordinate=self.shift_coordinate,
            use_udp=self.use_udp,
            dark_udp_refine=self.dark_udp_refine,
            kernel_size=self.kernel_size,
            flip_index=flip_index,
            shift_heatmap=shift_heatmap,
            dark_kernel=None if self.gaussian_kernel is None else self.gaussian_kernel[0, 0],
        )

    def construct(self, heatmap, center, scale, score) -> Tuple[torch.Tensor, torch.Tensor]:
        _, k, h, w = heatmap.shape
        return codec.topdown_decode(heatmap, center, scale, score, params=self._params(k, h, w))

    def decode_flip_pair(self, heatmap, flipped_heatmap, flip_index, center, scale, score,
                         shift_heatmap: bool = False):
        """Fused flip test: ``decoder((heatmap + flip_back(flipped)) * 0.5, ...)``
        without materialising the average (topdown_inferencer.py:165-187)."""
        _, k, h, w = heatmap.shape
        p = self._params(k, h, w, flip_index=flip_index, shift_heatmap=shift_heatmap)
        return codec.topdown_decode(heatmap, center, scale, score, flipped=flipped_heatmap,
                                    params=p)


@register("decoder", extra_name="bottomup_heatmap_ae")
class BottomUpHeatMapAEDecoder(Decoder):
    """Decode HigherHRNet heatmaps + associative-embedding tags into the top-M
    candidates per joint.

    Mirrors ``BottomUpHeatMapAEDecoder``
    (mindpose/models/decoders/bottom_up_decoder.py:13-203).

    Inputs: model_output = [out0 [N,2K,H0,W0], out1 [N,K,H1,W1]], mask [N,Hm,Wm]
    Outputs: val_k [N,K,M], tag_k [N,K,M,1], ind_k [N,K,M,2],
        heatmap_raw [N,K,H1,W1], tagging_heatmap [N,K,H1,W1,1]
    """

    def __init__(
        self,
        num_joints: int = 17,
        num_stages: int = 2,
        with_ae_loss: Sequence[bool] = (True, False),
        use_nms: bool = False,
        nms_kernel: int = 5,
        max_num: int = 30,
        tag_per_joint: bool = True,
        shift_coordinate: bool = False,
    ) -> None:
        self.num_joints = num_joints
        self.num_stages = num_stages
        self.with_ae_loss = list(with_ae_loss)
        self.use_nms = use_nms
        self.nms_kernel = nms_kernel
        self.max_num = max_num
        self.tag_per_joint = tag_per_joint
        self.shift_coordinate = shift_coordinate
        # materialise heatmap_raw / tagging_heatmap (only _refine_missing and the
        # visualiser read them); switch off to keep the decode at one read pass
        self.return_maps = True

    def construct(self, model_output: List[torch.Tensor], mask: torch.Tensor):
        heatmap, tagging_heatmap = self.decouple_output(model_output)
        return self.decode(heatmap, tagging_heatmap, mask, _raw=model_output)

    def decouple_output(self, output: List[torch.Tensor]):
        """Split the network output into heatmap / tag views
        (bottom_up_decoder.py:93-100); views, no copy."""
        heatmap, tagging_heatmap = list(), list()
        for i in range(self.num_stages):
            heatmap.append(output[i][:, : self.num_joints])
            if self.with_ae_loss[i]:
                tagging_heatmap.append(output[i][:, self.num_joints:])
        return heatmap, tagging_heatmap

    def decode(self, heatmap, tagging_heatmap, mask, _raw: Optional[List[torch.Tensor]] = None):
        from . import bottomup  # local: keeps top-down users free of it

        return bottomup.decode(self, heatmap, tagging_heatmap, mask, _raw)
