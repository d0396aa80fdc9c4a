"""Import the reference's numpy/cv2/scipy modules UNCHANGED from /root/reference.

Only usable in the build container (``/root/reference`` does not exist on the
GPU box).  Used by ``oracle/gen_golden.py`` and by the ``needs_reference`` tests
to pin the restatements in this package.

``mindpose/__init__.py`` and ``mindpose/data/__init__.py`` import MindSpore,
which is not installed, so empty parent packages whose ``__path__`` points into
the reference tree are registered first; the leaf modules (pure numpy / cv2 /
scipy) then import normally.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MINDPOSE_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mindpose"))


def _skeleton(name: str, rel: str) -> None:
    if name in sys.modules:
        return
    mod = types.ModuleType(name)
    mod.__path__ = [os.path.join(REFERENCE_ROOT, rel)]
    sys.modules[name] = mod


def load():
    """Returns a namespace with the reference modules of the numpy half."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _skeleton("mindpose", "mindpose")
    _skeleton("mindpose.data", "mindpose/data")
    _skeleton("mindpose.data.transform", "mindpose/data/transform")
    _skeleton("mindpose.utils", "mindpose/utils")
    ns = types.SimpleNamespace()
    ns.register = importlib.import_module("mindpose.register")
    ns.utils = importlib.import_module("mindpose.data.transform.utils")
    ns.topdown = importlib.import_module("mindpose.data.transform.topdown_transform")
    ns.bottomup = importlib.import_module("mindpose.data.transform.bottomup_transform")
    ns.match = importlib.import_module("mindpose.utils.match")
    ns.nms = importlib.import_module("mindpose.utils.nms")
    return ns
