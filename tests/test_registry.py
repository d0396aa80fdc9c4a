"""Registry and constructor contracts of the drop-in classes (CPU)."""
import logging

import numpy as np
import pytest

import mindpose_b200 as mp
from mindpose_b200 import register as reg
from mindpose_b200 import synth


def test_names_registered_like_the_reference():
    assert set(mp.list_modules()) >= {"decoder", "inferencer", "transform"}
    for name in ("bottomup_generate_target", "BottomUpGenerateTarget"):
        assert name in mp.list_components("transform")
    for name in ("topdown_box_to_center_scale", "topdown_affine", "topdown_generate_target",
                 "TopDownBoxToCenterScale", "TopDownAffine", "TopDownGenerateTarget"):
        assert name in mp.list_components("transform")
    for name in ("topdown_heatmap", "TopDownHeatMapDecoder", "bottomup_heatmap_ae",
                 "BottomUpHeatMapAEDecoder"):
        assert name in mp.list_components("decoder")
    for name in ("topdown_heatmap", "bottomup_heatmap_ae"):
        assert name in mp.list_components("inferencer")
    assert mp.entrypoint("decoder", "topdown_heatmap") is mp.entrypoint("decoder", "TopDownHeatMapDecoder")


def test_entrypoint_errors():
    with pytest.raises(ValueError, match="Unkown module"):
        mp.entrypoint("nope", "x")
    with pytest.raises(ValueError, match="Unkown components"):
        mp.entrypoint("decoder", "nope")


def test_duplicate_registration_warns_and_overrides(caplog):
    @reg.register("scratch", extra_name="alias")
    def first():
        return 1

    with caplog.at_level(logging.WARNING):
        @reg.register("scratch", extra_name="alias")
        def first():  # noqa: F811
            return 2

    assert "already registered" in caplog.text
    assert reg.entrypoint("scratch", "alias")() == 2
    assert reg.list_components("scratch") == ["alias", "first"]


@pytest.mark.needs_reference
def test_registry_names_are_a_subset_of_the_reference():
    from oracle import ref_loader

    ns = ref_loader.load()
    ref_names = set(ns.register.list_components("transform"))
    assert set(mp.list_components("transform")) <= ref_names


def test_decoder_constructor_contract():
    d = mp.create_decoder("topdown_heatmap", shift_coordinate=False)
    assert d.pixel_std == 200.0 and d.to_original and d.gaussian_kernel is None
    with pytest.raises(ValueError, match="cannot be `true` in the same time"):
        mp.create_decoder("topdown_heatmap", shift_coordinate=True, dark_udp_refine=True)
    d = mp.create_decoder("topdown_heatmap", dark_udp_refine=True, use_udp=True)
    k = d.gaussian_kernel
    assert k.shape == (1, 1, 11, 11) and k.dtype == np.float32 and abs(k.sum() - 1) < 1e-6
    from oracle import topdown_decode

    assert np.array_equal(k[0, 0], topdown_decode.dark_gaussian_kernel(11))
    b = mp.create_decoder("bottomup_heatmap_ae", use_nms=True, nms_kernel=3)
    assert b.max_num == 30 and b.num_stages == 2 and b.with_ae_loss == [True, False]


def test_transform_constructor_contract():
    cfg = dict(synth.TOPDOWN_CONFIG)
    t = mp.create_transform("topdown_generate_target", is_train=True, config=cfg, sigma=2.0)
    assert t._required_field[:3] == ["image", "center", "scale"] and len(t._required_field) == 8
    assert t._transform_cfg["flip_index"].tolist() == [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]
    with pytest.raises(ValueError, match="joint_weights"):
        mp.create_transform("topdown_generate_target", config=cfg, use_different_joint_weights=True)
    v = mp.create_transform("topdown_affine", is_train=False, config=cfg, use_udp=True)
    assert v._required_field[3] == "rotation" and v.use_udp


def test_inferencer_constructor_contract():
    cfg = dict(has_heatmap_output=True, hflip_tta=True, shift_heatmap=True,
               flip_pairs=synth.COCO_FLIP_PAIRS)
    with pytest.raises(ValueError, match="Decoder must be provided"):
        mp.create_inferencer(lambda x: x, "topdown_heatmap", config=cfg)
    with pytest.raises(ValueError, match="flip TTA need heatmap output"):
        mp.create_inferencer(lambda x: x, "topdown_heatmap",
                             config=dict(cfg, has_heatmap_output=False),
                             decoder=mp.create_decoder("topdown_heatmap"))
    inf = mp.create_inferencer(lambda x: x, "topdown_heatmap", config=cfg,
                               decoder=mp.create_decoder("topdown_heatmap"))
    assert inf._multi_run_net.flip_index.tolist()[:3] == [0, 2, 1]
