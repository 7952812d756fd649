"""``optree.typing`` stand-in: only the two names torchopt/typing.py:36 imports."""
from __future__ import annotations

from typing import Any, Generic, TypeVar

_T = TypeVar("_T")

__all__ = ["PyTree", "PyTreeTypeVar"]


class PyTree(Generic[_T]):  # ``PyTree[T]`` must be subscriptable; the alias carries no run-time behaviour
    def __class_getitem__(cls, item: Any) -> Any:
        return Any


def PyTreeTypeVar(name: str, param: Any) -> Any:  # noqa: N802, ARG001
    return Any
