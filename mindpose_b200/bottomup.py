"""Bottom-up codec entry points (placeholder until the kernels land)."""


def decode(decoder, heatmap, tagging_heatmap, mask, raw=None):
    raise NotImplementedError("bottom-up decode kernel not built yet")


def group_by_tag(*args, **kwargs):
    raise NotImplementedError("grouping kernel not built yet")


def transform_keypoints(*args, **kwargs):
    raise NotImplementedError("grouping kernel not built yet")
