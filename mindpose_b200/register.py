"""Component registry: the plugin boundary the codec sits behind.

Same contract as the reference's ``mindpose/register.py:12-59``:

* ``register(module, extra_name="")`` is a decorator that files the object under
  ``obj.__name__`` and, when given, under ``extra_name`` as well (``:20-36``);
* registering a name twice logs a warning and the newer object wins (``:13-14``);
* ``entrypoint(module, name)`` returns the object or raises ``ValueError``
  listing what is known (``:49-59``; the message text, typos included, is the
  reference's, because callers and tests match on it);
* ``list_modules()`` / ``list_components(module)`` return sorted names.
"""
import logging
from typing import Any, Callable, Dict, List

Component = Callable[..., Any]


class _Registry:
    """module name -> {component name -> object}; one instance per process."""

    def __init__(self) -> None:
        self._tables: Dict[str, Dict[str, Component]] = {}

    def add(self, module: str, names, obj: Component) -> Component:
        table = self._tables.setdefault(module, {})
        for name in names:
            if name in table:
                logging.warning(f"`{name}` is already registered")
            table[name] = obj
        return obj

    def modules(self) -> List[str]:
        return sorted(self._tables)

    def components(self, module: str) -> List[str]:
        return sorted(self._tables.get(module, ()))

    def lookup(self, module: str, name: str) -> Component:
        table = self._tables.get(module)
        if table is None:
            raise ValueError(f"Unkown module `{module}`. Supported modules: {self.modules()}")
        try:
            return table[name]
        except KeyError:
            raise ValueError(f"Unkown components `{name}`. Supported componetns in "
                             f"`{module}`: {self.components(module)}") from None


_REGISTRY = _Registry()


def register(module_name: str, extra_name: str = "") -> Callable[[Component], Component]:
    names = lambda obj: [obj.__name__] + ([extra_name] if extra_name else [])  # noqa: E731
    return lambda obj: _REGISTRY.add(module_name, names(obj), obj)


def list_modules() -> List[str]:
    return _REGISTRY.modules()


def list_components(module: str) -> List[str]:
    return _REGISTRY.components(module)


def entrypoint(module_name: str, component_name: str) -> Component:
    return _REGISTRY.lookup(module_name, component_name)
