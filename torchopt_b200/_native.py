"""Loads the native module ``torchopt_b200._C`` (pybind11/torch shim over libb200adam.so).

There is deliberately NO fallback: if the CUDA extension has not been built the import fails loudly
(the reference silently falls back to a pure-Python op, accelerated_op/adam_op.py:27-30; this product does not).
"""
from __future__ import annotations

import os

import torch  # noqa: F401  (libtorch must be resident before the extension is loaded)

_PKG = os.path.dirname(os.path.abspath(__file__))

try:
    from torchopt_b200 import _C  # type: ignore[attr-defined]
except ImportError as exc:  # pragma: no cover - exercised only on a broken install
    raise ImportError(
        "torchopt_b200: the native CUDA extension is missing or failed to load "
        f"({exc}). Build it in-tree with `python -c 'import __graft_entry__ as g; g.build()'` "
        f"(expects {os.path.join(_PKG, 'libb200adam.so')} and {os.path.join(_PKG, '_C.so')}). "
        "There is no CPU or pure-Python fallback in this package."
    ) from exc

adam_op = _C.adam_op

__all__ = ["_C", "adam_op"]
