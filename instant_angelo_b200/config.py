"""Config loading for the reference's YAML files without omegaconf (not installed in this image).

Mirrors reference utils/misc.py:7-31: `${a.b}` interpolation, the resolvers `add sub mul div idiv
basename calc_exp_lr_decay_rate`, and CLI dot-list overrides (`model.geometry.grad_type=finite_difference`).
"""
from __future__ import annotations

import os
import re
from typing import Any, Iterable

import yaml


class Config(dict):
    """dict with attribute access, like an OmegaConf DictConfig for the purposes of the hot path."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def copy(self):
        return Config(super().copy())


def to_config(obj: Any) -> Any:
    if isinstance(obj, dict):
        return Config({k: to_config(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return [to_config(v) for v in obj]
    return obj


def to_primitive(obj: Any) -> Any:
    if isinstance(obj, dict):
        return {k: to_primitive(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [to_primitive(v) for v in obj]
    return obj


_RESOLVERS = {
    "add": lambda a, b: a + b,
    "sub": lambda a, b: a - b,
    "mul": lambda a, b: a * b,
    "div": lambda a, b: a / b,
    "idiv": lambda a, b: a // b,
    "basename": lambda p: os.path.basename(str(p)),
    "calc_exp_lr_decay_rate": lambda factor, n: factor ** (1.0 / n),
}


def _parse_scalar(s: str) -> Any:
    try:
        return yaml.safe_load(s)
    except yaml.YAMLError:
        return s


def _split_args(s: str) -> list:
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
            continue
        depth += ch == "{"
        depth -= ch == "}"
        cur += ch
    out.append(cur)
    return [a.strip() for a in out]


_INNER = re.compile(r"\$\{([^${}]*)\}")


def _lookup(root: dict, path: str) -> Any:
    node: Any = root
    for part in path.split("."):
        node = node[int(part)] if isinstance(node, list) else node[part]
    return node


def _resolve_str(root: dict, s: str, depth: int = 0) -> Any:
    if depth > 32:
        raise ValueError(f"interpolation too deep: {s}")
    while True:
        m = _INNER.search(s)
        if m is None:
            return s
        expr = m.group(1)
        if ":" in expr and expr.split(":", 1)[0] in _RESOLVERS:
            name, argstr = expr.split(":", 1)
            args = [_parse_scalar(a) if not isinstance(a, (int, float)) else a for a in _split_args(argstr)]
            val = _RESOLVERS[name](*args)
        else:
            val = _lookup(root, expr)
            if val == "???":
                val = "???"
            if isinstance(val, str) and "${" in val:
                val = _resolve_str(root, val, depth + 1)
        if m.start() == 0 and m.end() == len(s):
            return val
        s = s[: m.start()] + str(val) + s[m.end():]


def _resolve_tree(root: dict, node: Any) -> Any:
    if isinstance(node, dict):
        for k in list(node.keys()):
            node[k] = _resolve_tree(root, node[k])
        return node
    if isinstance(node, list):
        return [_resolve_tree(root, v) for v in node]
    if isinstance(node, str) and "${" in node:
        try:
            return _resolve_str(root, node)
        except (KeyError, TypeError):
            return node
    return node


def apply_overrides(cfg: dict, overrides: Iterable[str]) -> None:
    for item in overrides:
        key, val = item.split("=", 1)
        node = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = _parse_scalar(val)


def load_config(*yaml_files: str, cli_args: Iterable[str] = ()) -> Config:
    merged: dict = {}

    def merge(dst, src):
        for k, v in src.items():
            if isinstance(v, dict) and isinstance(dst.get(k), dict):
                merge(dst[k], v)
            else:
                dst[k] = v

    for f in yaml_files:
        with open(f) as fp:
            merge(merged, yaml.safe_load(fp))
    apply_overrides(merged, cli_args)
    _resolve_tree(merged, merged)
    return to_config(merged)
