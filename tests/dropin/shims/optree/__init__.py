"""Test-only stand-in for the third-party ``optree`` package (pure Python).

TorchOpt imports ``optree`` at package import (torchopt/pytree.py:23-27, torchopt/typing.py:36); it is not
installed in this image and does no arithmetic, so the drop-in acceptance harness (tests/test_dropin_gpu.py)
puts this module on PYTHONPATH instead.  Implements the subset TorchOpt's Adam path uses, with optree's
conventions: dict keys are visited in sorted order, ``None`` is an internal node without children unless
``none_is_leaf=True``, namedtuples keep their type.  NOT part of the product.
"""
from __future__ import annotations

from collections import OrderedDict, defaultdict, deque
from typing import Any, Callable, Iterable

from optree import typing  # noqa: F401  (optree.typing must be importable as a submodule)

__all__ = ["PyTreeSpec", "tree_flatten", "tree_unflatten", "tree_leaves", "tree_structure", "tree_map", "tree_map_",
           "tree_transpose", "tree_all", "tree_any"]


class PyTreeSpec:
    __slots__ = ("kind", "meta", "children", "num_leaves", "none_is_leaf")

    def __init__(self, kind: str, meta: Any, children: tuple, none_is_leaf: bool) -> None:
        self.kind, self.meta, self.children, self.none_is_leaf = kind, meta, children, none_is_leaf
        self.num_leaves = 1 if kind == "leaf" else sum(c.num_leaves for c in children)

    @property
    def num_nodes(self) -> int:
        return 1 + sum(c.num_nodes for c in self.children)

    @property
    def num_children(self) -> int:
        return len(self.children)

    def _key(self) -> tuple:
        return (self.kind, self.meta, tuple(c._key() for c in self.children))

    def __eq__(self, other: Any) -> bool:
        return isinstance(other, PyTreeSpec) and self._key() == other._key()

    def __ne__(self, other: Any) -> bool:
        return not self == other

    def __hash__(self) -> int:
        return hash(repr(self._key()))

    def __repr__(self) -> str:
        return f"PyTreeSpec({self.kind}, num_leaves={self.num_leaves})"

    def unflatten(self, leaves: Iterable) -> Any:
        it = iter(leaves)
        out = _unflatten(self, it)
        return out

    def flatten_up_to(self, tree: Any) -> list:
        out: list = []
        _flatten_up_to(self, tree, out)
        return out


def _is_namedtuple(x: Any) -> bool:
    return isinstance(x, tuple) and hasattr(type(x), "_fields")


def _flatten(node: Any, leaves: list, is_leaf: Callable | None, nil: bool) -> PyTreeSpec:
    if is_leaf is not None and is_leaf(node):
        leaves.append(node)
        return PyTreeSpec("leaf", None, (), nil)
    if node is None:
        if nil:
            leaves.append(node)
            return PyTreeSpec("leaf", None, (), nil)
        return PyTreeSpec("none", None, (), nil)
    if _is_namedtuple(node):
        return PyTreeSpec("namedtuple", type(node), tuple(_flatten(c, leaves, is_leaf, nil) for c in node), nil)
    if isinstance(node, tuple):
        return PyTreeSpec("tuple", None, tuple(_flatten(c, leaves, is_leaf, nil) for c in node), nil)
    if isinstance(node, list):
        return PyTreeSpec("list", None, tuple(_flatten(c, leaves, is_leaf, nil) for c in node), nil)
    if isinstance(node, deque):
        return PyTreeSpec("deque", node.maxlen, tuple(_flatten(c, leaves, is_leaf, nil) for c in node), nil)
    if isinstance(node, OrderedDict):
        keys = tuple(node)
        return PyTreeSpec("ordereddict", keys, tuple(_flatten(node[k], leaves, is_leaf, nil) for k in keys), nil)
    if isinstance(node, defaultdict):
        keys = tuple(sorted(node))
        return PyTreeSpec("defaultdict", (node.default_factory, keys),
                          tuple(_flatten(node[k], leaves, is_leaf, nil) for k in keys), nil)
    if isinstance(node, dict):
        keys = tuple(sorted(node))
        return PyTreeSpec("dict", keys, tuple(_flatten(node[k], leaves, is_leaf, nil) for k in keys), nil)
    leaves.append(node)
    return PyTreeSpec("leaf", None, (), nil)


def _build(spec: PyTreeSpec, kids: list) -> Any:
    k = spec.kind
    if k == "none":
        return None
    if k == "tuple":
        return tuple(kids)
    if k == "list":
        return list(kids)
    if k == "namedtuple":
        return spec.meta(*kids)
    if k == "deque":
        return deque(kids, maxlen=spec.meta)
    if k == "ordereddict":
        return OrderedDict(zip(spec.meta, kids))
    if k == "defaultdict":
        return defaultdict(spec.meta[0], zip(spec.meta[1], kids))
    return dict(zip(spec.meta, kids))


def _unflatten(spec: PyTreeSpec, it: Any) -> Any:
    if spec.kind == "leaf":
        return next(it)
    return _build(spec, [_unflatten(c, it) for c in spec.children])


def _children_of(spec: PyTreeSpec, node: Any) -> list:
    k = spec.kind
    if k == "none":
        if node is not None:
            raise ValueError(f"Expected None, got {node!r}.")
        return []
    if k in ("tuple", "list", "namedtuple", "deque"):
        if not isinstance(node, (tuple, list, deque)) or len(node) != len(spec.children):
            raise ValueError(f"pytree structure mismatch at {node!r}")
        return list(node)
    keys = spec.meta[1] if k == "defaultdict" else spec.meta
    if not isinstance(node, dict) or set(node) != set(keys):
        raise ValueError(f"pytree structure mismatch at {node!r}")
    return [node[key] for key in keys]


def _flatten_up_to(spec: PyTreeSpec, node: Any, out: list) -> None:
    if spec.kind == "leaf":
        out.append(node)
        return
    for c, sub in zip(spec.children, _children_of(spec, node)):
        _flatten_up_to(c, sub, out)


def tree_flatten(tree: Any, is_leaf: Callable | None = None, *, none_is_leaf: bool = False,
                 namespace: str = "") -> tuple[list, PyTreeSpec]:  # noqa: ARG001
    leaves: list = []
    spec = _flatten(tree, leaves, is_leaf, none_is_leaf)
    return leaves, spec


def tree_unflatten(treespec: PyTreeSpec, leaves: Iterable) -> Any:
    return treespec.unflatten(leaves)


def tree_leaves(tree: Any, is_leaf: Callable | None = None, *, none_is_leaf: bool = False,
                namespace: str = "") -> list:  # noqa: ARG001
    return tree_flatten(tree, is_leaf, none_is_leaf=none_is_leaf)[0]


def tree_structure(tree: Any, is_leaf: Callable | None = None, *, none_is_leaf: bool = False,
                   namespace: str = "") -> PyTreeSpec:  # noqa: ARG001
    return tree_flatten(tree, is_leaf, none_is_leaf=none_is_leaf)[1]


def tree_map(func: Callable, tree: Any, *rests: Any, is_leaf: Callable | None = None, none_is_leaf: bool = False,
             namespace: str = "") -> Any:  # noqa: ARG001
    leaves, spec = tree_flatten(tree, is_leaf, none_is_leaf=none_is_leaf)
    cols = [leaves] + [spec.flatten_up_to(r) for r in rests]
    return spec.unflatten(func(*xs) for xs in zip(*cols))


def tree_map_(func: Callable, tree: Any, *rests: Any, is_leaf: Callable | None = None, none_is_leaf: bool = False,
              namespace: str = "") -> Any:  # noqa: ARG001
    leaves, spec = tree_flatten(tree, is_leaf, none_is_leaf=none_is_leaf)
    cols = [leaves] + [spec.flatten_up_to(r) for r in rests]
    for xs in zip(*cols):
        func(*xs)
    return tree


def tree_transpose(outer_treespec: PyTreeSpec, inner_treespec: PyTreeSpec, tree: Any) -> Any:
    """``outer[inner[x]] -> inner[outer[x]]``."""
    per_outer = [inner_treespec.flatten_up_to(sub) for sub in outer_treespec.flatten_up_to(tree)]
    n_inner = inner_treespec.num_leaves
    cols = [outer_treespec.unflatten(row[j] for row in per_outer) for j in range(n_inner)]
    return inner_treespec.unflatten(cols)


def tree_all(tree: Any, is_leaf: Callable | None = None, *, none_is_leaf: bool = False, namespace: str = "") -> bool:
    return all(tree_leaves(tree, is_leaf, none_is_leaf=none_is_leaf))


def tree_any(tree: Any, is_leaf: Callable | None = None, *, none_is_leaf: bool = False, namespace: str = "") -> bool:
    return any(tree_leaves(tree, is_leaf, none_is_leaf=none_is_leaf))
