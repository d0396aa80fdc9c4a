"""Host-side behaviour of the N4 mirror (mindpose_b200/nms.py) that needs no GPU: the
reference's empty-input convention and the no-CPU-fallback rule."""
import numpy as np
import pytest
import torch

from mindpose_b200 import nms as dnms


def test_empty_input_returns_empty_list_like_the_reference():
    # mindpose/utils/nms.py:87-88 / :160-161: `if not kpts_db: return []`
    assert dnms.oks_nms([], 0.9) == []
    assert dnms.soft_oks_nms([], 0.9) == []


def test_no_cpu_fallback():
    kp = torch.zeros(3, 17, 3)
    ar = torch.ones(3)
    sc = torch.ones(3)
    off = torch.tensor([0, 3], dtype=torch.int32)
    with pytest.raises(ValueError, match="CUDA"):
        dnms.rescore_and_nms(kp, ar, sc, off, 3, oks_thr=0.9)


def test_sigmas_default_is_the_reference_default():
    want = np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07, .87, .87,
                     .89, .89]) / 10.0
    assert np.array_equal(dnms.COCO_SIGMAS, want)
