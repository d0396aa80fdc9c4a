"""Column order of the dataset pipeline (mindpose/data/column_names.py:4-86).

``Transform.__call__`` zips positional columns against these names, so the order
is part of the calling convention the drop-in transforms keep.
"""

_TOPDOWN = {
    "train": "image center scale boxes keypoints rotation target target_weight".split(),
    "val": "image center scale rotation image_file boxes bbox_ids bbox_scores".split(),
}
_TOPDOWN_FINAL = {
    "train": "image target target_weight".split(),
    "val": "image image_file boxes bbox_ids center scale bbox_scores".split(),
}
_BOTTOMUP = {
    "train": "image boxes keypoints target mask tag_ind".split(),
    "val": "image mask center scale image_file image_shape".split(),
}
_BOTTOMUP_FINAL = {
    "train": "image target mask tag_ind".split(),
    "val": "image mask center scale image_file image_shape".split(),
}

COLUMN_MAP = {
    "coco_topdown": _TOPDOWN,
    "topdown": _TOPDOWN,
    "coco_bottomup": _BOTTOMUP,
    "bottomup": _BOTTOMUP,
    "imagefolder_bottomup": {"val": _BOTTOMUP["val"]},
}

FINAL_COLUMN_MAP = {
    "topdown": _TOPDOWN_FINAL,
    "bottomup": _BOTTOMUP_FINAL,
    "imagefolder_bottomup": {"val": _BOTTOMUP_FINAL["val"]},
}
