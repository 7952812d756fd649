"""``scale_by_accelerated_adam`` -- the caller of the hot path, re-built around ONE multi-tensor launch.

Mirrors /root/reference/torchopt/transform/scale_by_adam.py:225-420 (same constructor arguments, same
``GradientTransformation(init, update)`` protocol of torchopt/base.py:118-151, same
``ScaleByAdamState(mu, nu, count)`` state with ``count`` a per-leaf 0-dim int64 tensor on the leaf's device),
but ``update`` hands every leaf to ``AdamOp.multi`` at once instead of looping ``op(mu, nu, g, count)`` in Python,
and the step counters carry a host mirror:

* reference, per leaf per step: ``count + (count != INT64_MAX)`` = 3 tiny kernels (transform/utils.py:117-121)
  plus a device->host sync when pybind converts the count tensor to ``size_t``;
* here, per step: one small host->device copy builds all new counters; the kernels get their bias corrections
  from the host mirror (``tensor._b200_host_count``), so nothing synchronises.  A state whose counters lack the
  mirror (e.g. one deserialised by the user) is read once with ``int()`` and tagged.
"""
from __future__ import annotations

from typing import Any, Callable, NamedTuple, Sequence

import torch

from torchopt_b200 import _pytree
from torchopt_b200.accelerated_op.adam_op import (INT64_MAX, AdamOp, capturing, check_capturable, count_as_int,
                                                    tag_count)


# A/B switch for parity tests: False routes the differentiable path through the reference's three-node graph
# (MuOp/NuOp/UpdatesOp, three launches per leaf) instead of the fused single node.
FUSED_GRAPH = True


class GradientTransformation(NamedTuple):
    """``(init, update)`` pair -- torchopt/base.py:118-151."""

    init: Callable
    update: Callable


class ScaleByAdamState(NamedTuple):
    """State for the Adam algorithm -- transform/scale_by_adam.py:58-63."""

    mu: Any
    nu: Any
    count: Any


# dtype of the Adam moments for a given parameter dtype (identity unless listed)
STATE_DTYPE = {torch.bfloat16: torch.float32}


_tag = tag_count


def make_counts(values: Sequence[int | None], devices: Sequence[torch.device | None]) -> list:
    """0-dim int64 device tensors holding ``values`` (views of ONE vector per device), each tagged with its host
    mirror.  ``None`` entries stay ``None``.  When all counters of a device agree -- the normal case -- the vector is
    produced by one fill kernel; otherwise by one small host->device copy."""
    out: list = [None] * len(values)
    by_dev: dict = {}
    in_capture = capturing()
    for i, (v, d) in enumerate(zip(values, devices)):
        if v is not None:
            by_dev.setdefault(d, []).append(i)
    for dev, idx in by_dev.items():
        vals = [values[i] for i in idx]
        if all(v == vals[0] for v in vals):
            vec = torch.full((len(idx),), vals[0], dtype=torch.int64, device=dev)
        elif in_capture:
            # a host->device copy from a temporary pinned buffer must not be recorded (the buffer is gone at replay):
            # one fill per distinct value instead
            vec = torch.empty((len(idx),), dtype=torch.int64, device=dev)
            for v in sorted(set(vals)):
                sel = [k for k, x in enumerate(vals) if x == v]
                lo, hi = sel[0], sel[-1] + 1
                if sel == list(range(lo, hi)):
                    vec[lo:hi].fill_(v)
                else:
                    for k in sel:
                        vec[k].fill_(v)
        else:
            host = torch.tensor(vals, dtype=torch.int64)
            if dev.type == "cuda" and torch.cuda.is_available():
                vec = host.pin_memory().to(dev, non_blocking=True)
            else:
                vec = host.to(dev)
        for i, view in zip(idx, vec.unbind(0)):
            out[i] = _tag(view, values[i], in_capture)
    return out


def zeros_like_flat(leaves: Sequence, requires_grad: bool = False, dtype_map: dict | None = None) -> list:
    """``torch.zeros_like(p, requires_grad=...)`` (optionally with a remapped dtype) for every non-None leaf, carved out of ONE zero-filled buffer per
    (device, dtype): one allocation + one memset instead of one per leaf.  Every result has its
    leaf's sizes and -- for dense leaves -- strides (what ``zeros_like`` preserves), starts on a 128-byte boundary
    and is an independent autograd leaf."""
    out: list = [None] * len(leaves)
    groups: dict = {}
    dtype_map = dtype_map or {}
    for i, p in enumerate(leaves):
        if p is not None:
            groups.setdefault((p.device, dtype_map.get(p.dtype, p.dtype)), []).append(i)
    for (dev, dtype), idx in groups.items():
        if len(idx) == 1:
            out[idx[0]] = torch.zeros_like(leaves[idx[0]], dtype=dtype, requires_grad=requires_grad)
            continue
        align = max(1, 128 // torch.empty((), dtype=dtype).element_size())
        offsets, total = [], 0
        for i in idx:
            total = (total + align - 1) // align * align
            offsets.append(total)
            total += leaves[i].numel()
        flat = torch.zeros(total, dtype=dtype, device=dev)
        for i, off in zip(idx, offsets):
            p = leaves[i]
            # zeros_like keeps the strides of dense leaves (e.g. channels_last / transposed); ask empty_like in the
            # rare non-contiguous case rather than re-deriving ATen's rule
            strides = p.stride() if p.is_contiguous() else torch.empty_like(p).stride()
            t = flat.as_strided(p.shape, strides, off)
            out[i] = t.requires_grad_(True) if requires_grad else t
    return out


def inc_count_flat(updates: Sequence, counts: Sequence) -> list:
    """``count + 1`` (saturating at INT64_MAX) for leaves that received an update; others pass through.
    Same values as transform/utils.py:111-122, built from the host mirrors."""
    values, devices = [], []
    checked = False
    for g, c in zip(updates, counts):
        if g is None or c is None:
            values.append(None), devices.append(None)
            continue
        if not checked:          # one counter speaks for the state (they are created together)
            check_capturable(c)
            checked = True
        host = count_as_int(c)   # tags an untagged / edited counter after reading it once
        values.append(host + (host != INT64_MAX))
        devices.append(c.device if isinstance(c, torch.Tensor) else torch.device("cpu"))
    fresh = make_counts(values, devices)
    return [fresh[i] if values[i] is not None else counts[i] for i in range(len(counts))]


# pylint: disable-next=too-many-arguments
def _scale_by_accelerated_adam(b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8, eps_root: float = 0.0,
                               moment_requires_grad: bool = False, *,
                               already_flattened: bool = False) -> GradientTransformation:
    # argument validation as in transform/scale_by_adam.py:292-299
    if not eps >= 0.0:
        raise ValueError(f"Invalid epsilon value: {eps}")
    if not 0.0 <= b1 < 1.0:
        raise ValueError(f"Invalid beta parameter at index 0: {b1}")
    if not 0.0 <= b2 < 1.0:
        raise ValueError(f"Invalid beta parameter at index 1: {b2}")

    def init_flat(params: Sequence) -> ScaleByAdamState:
        leaves = list(params)
        count = make_counts([None if p is None else 0 for p in leaves],
                            [None if p is None else p.device for p in leaves])
        # bf16 parameters get fp32 moments (the "bf16-param / fp32-state" variant): the reference would create bf16
        # moments (zeros_like), in which b2 = 0.999 rounds to 1.0 and nu never accumulates
        with torch.no_grad():
            mu = zeros_like_flat(leaves, requires_grad=moment_requires_grad, dtype_map=STATE_DTYPE)
            nu = zeros_like_flat(leaves, requires_grad=moment_requires_grad, dtype_map=STATE_DTYPE)
        return ScaleByAdamState(mu=mu, nu=nu, count=count)

    def update_flat(updates: Sequence, state: ScaleByAdamState, inplace: bool):
        op = AdamOp(b1=b1, b2=b2, eps=eps, eps_root=eps_root, inplace=inplace, fused=FUSED_GRAPH)
        count_inc = inc_count_flat(updates, state.count)
        new_mu, new_nu, new_updates = op.multi(state.mu, state.nu, updates, count_inc)
        return new_updates, new_mu, new_nu, count_inc

    def rewrap(new: list, like: Any) -> Any:  # keep the container type of the caller (list / tuple)
        return new if isinstance(like, list) else type(like)(new) if isinstance(like, tuple) else new

    if already_flattened:

        def init_fn(params: Sequence) -> ScaleByAdamState:
            st = init_flat(params)
            return ScaleByAdamState(*(rewrap(x, params) for x in st))

        def update_fn(updates: Sequence, state: ScaleByAdamState, *, params: Any = None,  # noqa: ARG001
                      inplace: bool = True):
            new_updates, new_mu, new_nu, count_inc = update_flat(updates, state, inplace)
            return rewrap(new_updates, updates), ScaleByAdamState(
                mu=rewrap(new_mu, state.mu), nu=rewrap(new_nu, state.nu), count=rewrap(count_inc, state.count))

    else:

        def init_fn(params: Any) -> ScaleByAdamState:
            leaves, spec = _pytree.tree_flatten(params)
            st = init_flat(leaves)
            return ScaleByAdamState(*(_pytree.tree_unflatten(spec, x) for x in st))

        def update_fn(updates: Any, state: ScaleByAdamState, *, params: Any = None,  # noqa: ARG001
                      inplace: bool = True):
            leaves, spec = _pytree.tree_flatten(updates)
            flat_state = ScaleByAdamState(*(_pytree.tree_flatten(x)[0] for x in state))
            new_updates, new_mu, new_nu, count_inc = update_flat(leaves, flat_state, inplace)
            un = _pytree.tree_unflatten
            return un(spec, new_updates), ScaleByAdamState(mu=un(spec, new_mu), nu=un(spec, new_nu),
                                                           count=un(spec, count_inc))

    return GradientTransformation(init_fn, update_fn)


def scale_by_accelerated_adam(b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8, eps_root: float = 0.0,
                              moment_requires_grad: bool = False) -> GradientTransformation:
    """Rescale updates according to the Adam algorithm with the B200 fused operator (pytree variant)."""
    return _scale_by_accelerated_adam(b1, b2, eps, eps_root, moment_requires_grad, already_flattened=False)


def _flat(b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8, eps_root: float = 0.0,
          moment_requires_grad: bool = False) -> GradientTransformation:
    return _scale_by_accelerated_adam(b1, b2, eps, eps_root, moment_requires_grad, already_flattened=True)


scale_by_accelerated_adam.flat = _flat  # type: ignore[attr-defined]
scale_by_accelerated_adam.impl = _scale_by_accelerated_adam  # type: ignore[attr-defined]
