"""A minimal pytree flatten/unflatten with ``none_is_leaf=True`` semantics.

The reference delegates this to the third-party ``optree`` package (torchopt/pytree.py:23-27), which is not
part of the hot path and does no arithmetic.  Only what ``scale_by_accelerated_adam`` needs is implemented:
tuples, lists, dicts (sorted keys, as optree does), namedtuples; everything else -- including ``None`` -- is a leaf.
"""
from __future__ import annotations

from typing import Any, Callable, NamedTuple


class TreeSpec(NamedTuple):
    kind: str            # 'leaf' | 'tuple' | 'list' | 'dict' | 'namedtuple'
    meta: Any            # dict keys / namedtuple type
    children: tuple
    num_leaves: int


_LEAF = TreeSpec("leaf", None, (), 1)


def _is_namedtuple(x: Any) -> bool:
    return isinstance(x, tuple) and hasattr(type(x), "_fields")


def _flatten_into(node: Any, leaves: list) -> TreeSpec:
    # a module-level function on purpose: a nested closure that calls itself forms a reference cycle
    # (function -> cell -> function) that would keep ``leaves`` -- tensors with their autograd graphs -- alive until the
    # next garbage-collection pass
    if _is_namedtuple(node):
        kids = tuple(_flatten_into(c, leaves) for c in node)
        return TreeSpec("namedtuple", type(node), kids, sum(k.num_leaves for k in kids))
    if isinstance(node, (tuple, list)):
        kids = tuple(_flatten_into(c, leaves) for c in node)
        return TreeSpec("tuple" if isinstance(node, tuple) else "list", None, kids, sum(k.num_leaves for k in kids))
    if isinstance(node, dict):
        keys = sorted(node)
        kids = tuple(_flatten_into(node[k], leaves) for k in keys)
        return TreeSpec("dict", (type(node), tuple(keys)), kids, sum(k.num_leaves for k in kids))
    leaves.append(node)
    return _LEAF


def tree_flatten(tree: Any) -> tuple[list, TreeSpec]:
    leaves: list = []
    spec = _flatten_into(tree, leaves)
    return leaves, spec


def _unflatten_from(s: TreeSpec, it: Any) -> Any:
    if s.kind == "leaf":
        return next(it)
    kids = [_unflatten_from(c, it) for c in s.children]
    if s.kind == "tuple":
        return tuple(kids)
    if s.kind == "list":
        return kids
    if s.kind == "namedtuple":
        return s.meta(*kids)
    typ, keys = s.meta
    return typ(zip(keys, kids)) if typ is not dict else dict(zip(keys, kids))


def tree_unflatten(spec: TreeSpec, leaves) -> Any:
    return _unflatten_from(spec, iter(leaves))


def tree_map(fn: Callable, tree: Any, *rests: Any) -> Any:
    leaves, spec = tree_flatten(tree)
    others = [tree_flatten(r)[0] for r in rests]
    for o in others:
        if len(o) != len(leaves):
            raise ValueError("tree_map: pytrees have different structures")
    return tree_unflatten(spec, [fn(*xs) for xs in zip(leaves, *others)])
