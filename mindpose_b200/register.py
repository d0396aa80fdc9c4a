"""Component registry: the plugin boundary the codec sits behind.

Same contract as the reference's ``mindpose/register.py:12-59``:

* ``register(module, extra_name="")`` is a decorator that files the object under
  ``obj.__name__`` and, when given, under ``extra_name`` as well (``:20-36``);
* registering a name twice logs a warning and the newer object wins (``:13-14``);
* ``entrypoint(module, name)`` returns the object or raises ``ValueError``
  listing what is known (``:49-59``);
* ``list_modules()`` / ``list_components(module)`` return sorted names.
"""
import logging
from typing import Any, Callable, Dict, List

_REGISTRY: Dict[str, Dict[str, Callable[..., Any]]] = {}


def _file(module_name: str, name: str, obj: Callable[..., Any]) -> None:
    table = _REGISTRY.setdefault(module_name, {})
    if name in table:
        logging.warning(f"`{name}` is already registered")
    table[name] = obj


def register(module_name: str, extra_name: str = "") -> Callable[..., Any]:
    def decorate(obj: Callable[..., Any]) -> Callable[..., Any]:
        _file(module_name, obj.__name__, obj)
        if extra_name:
            _file(module_name, extra_name, obj)
        return obj

    return decorate


def list_modules() -> List[str]:
    return sorted(_REGISTRY)


def list_components(module: str) -> List[str]:
    return sorted(_REGISTRY.get(module, {}))


def entrypoint(module_name: str, component_name: str) -> Callable[..., Any]:
    if module_name not in _REGISTRY:
        raise ValueError(
            f"Unkown module `{module_name}`. Supported modules: {list_modules()}"
        )
    table = _REGISTRY[module_name]
    if component_name not in table:
        raise ValueError(
            f"Unkown components `{component_name}`. "
            f"Supported componetns in `{module_name}`: {list_components(module_name)}"
        )
    return table[component_name]
