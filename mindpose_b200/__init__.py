"""mindpose_b200 -- sm_100a (B200) implementation of mindpose's heatmap codec
hot path behind mindpose's own registry names.

Importing the package registers the drop-in transforms, decoders and
inferencers (``register.entrypoint("decoder", "topdown_heatmap")`` ...).  The
CUDA library is loaded on first use; there is no CPU fallback.
"""
from . import register  # noqa: F401
from . import column_names  # noqa: F401
from . import decoders  # noqa: F401
from . import inferencers  # noqa: F401
from . import transforms  # noqa: F401
from .register import entrypoint, list_components, list_modules  # noqa: F401


def create_decoder(name: str, **kwargs):
    """mindpose/models/model_factory.py:82-92."""
    return entrypoint("decoder", name)(**kwargs)


def create_transform(name: str, is_train: bool = True, config=None, **kwargs):
    """How data_factory builds a transform (mindpose/data/data_factory.py:168-170)."""
    return entrypoint("transform", name)(is_train=is_train, config=config, **kwargs)


def create_inferencer(net, name: str, config=None, **kwargs):
    """mindpose/engine/factory.py:13-43 (config already merged)."""
    return entrypoint("inferencer", name)(net, config=config, **kwargs)


__version__ = "0.1.0"
