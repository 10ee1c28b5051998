"""Name -> class table behind `register(name)` / `make(name, config)`; the reference resolves its models the same way
(models/__init__.py:1-13), and the names registered here are the ones its configs use ('neus', 'volume-sdf', ...)."""
from __future__ import annotations

from typing import Callable, Dict, Type

_TABLE: Dict[str, Type] = {}
models = _TABLE          # the reference exposes the table under this name


def register(name: str) -> Callable[[Type], Type]:
    """Class decorator: make `cls` constructible as make(name, config).  Re-registering a name is an error here (the
    reference silently overwrites), because two kernels-backed classes under one name would make parity runs ambiguous."""

    def bind(cls: Type) -> Type:
        if name in _TABLE and _TABLE[name] is not cls:
            raise KeyError(f"model name {name!r} is already registered to {_TABLE[name].__name__}")
        _TABLE[name] = cls
        return cls

    return bind


def make(name: str, config):
    try:
        cls = _TABLE[name]
    except KeyError:
        raise KeyError(f"unknown model {name!r}; registered: {sorted(_TABLE)}") from None
    return cls(config)
