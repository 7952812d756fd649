"""Thin callers of the hot path: ``adam`` / ``adamw`` recipes, ``apply_updates``, ``MetaAdam``, ``Adam``.

These are the reference's *unchanged callers* ("OUT OF SCOPE -- unchanged callers"): in a real
deployment TorchOpt's own ``torchopt.adam`` / ``torchopt.MetaAdam`` sit here and only ``torchopt._C`` +
``torchopt/accelerated_op/adam_op.py`` are swapped.  They are restated minimally (flat lists of
tensors only) so the BASELINE configs -- ``adam(use_accelerated_op=True)`` K-step unrolls, ``MetaAdam`` MAML inner
loops -- can run on a box that has no TorchOpt installed.  Reference call stack: alias/adam.py:116-137,
optim/meta/base.py:58-96, optim/base.py:99-124, update.py:69-81, transform/scale.py:86-105.
"""
from __future__ import annotations

from typing import Any, Iterable, Sequence

import torch
from torch import nn

from torchopt_b200 import _pytree
from torchopt_b200._native import adam_op
from torchopt_b200.accelerated_op.adam_step import AdamStep
from torchopt_b200.accelerated_op.adam_op import INT64_MAX, capturing, count_as_int
from torchopt_b200.transform import (STATE_DTYPE, GradientTransformation, ScaleByAdamState, inc_count_flat,
                                     make_counts, scale_by_accelerated_adam)

# A/B switch for parity tests: False makes MetaAdam / Adam run the chained transformations + apply_updates
# (one Adam launch + 2 torch kernels per leaf per step) instead of the single whole-step launch.
FUSE_STEP = True
# A/B switch for parity tests: False keeps the optimizer state as per-leaf tensors (the reference's ScaleByAdamState
# layout; 4L autograd inputs / 3L outputs per step) instead of the default flat moment buffers (2L+2 / L+2).
FLAT_STATE = True
# Measurement hook (bench.py): when True, MetaAdam.step records in ``last_host_us`` how its host time splits between
# ``autograd.grad`` of the caller's loss and the optimizer update proper.
PROFILE_HOST = False


def chain_flat(*transforms: GradientTransformation) -> GradientTransformation:
    """Compose transformations over a flat list of leaves (combine.py:49-102, base.py:184-200)."""
    transforms = tuple(t for t in transforms if t is not None)

    def init_fn(params: Sequence) -> tuple:
        return tuple(t.init(params) for t in transforms)

    def update_fn(updates: Sequence, state: tuple, *, params: Any = None, inplace: bool = True):
        new_state = []
        for t, s in zip(transforms, state):
            updates, s = t.update(updates, s, params=params, inplace=inplace)
            new_state.append(s)
        return updates, tuple(new_state)

    return GradientTransformation(init_fn, update_fn)


def with_flattened_tree(inner: GradientTransformation) -> GradientTransformation:
    """Let a transformation written for flat leaf lists accept arbitrary pytrees (dicts, nested lists / tuples,
    namedtuples; ``None`` is a leaf) -- base.py:203-237 / combine.py:49-102: ``init`` and ``update`` flatten their
    arguments, the state stays flat, the updates come back in the structure they arrived in."""

    def init_fn(params: Any) -> Any:
        return inner.init(_pytree.tree_flatten(params)[0])

    def update_fn(updates: Any, state: Any, *, params: Any = None, inplace: bool = True):
        leaves, spec = _pytree.tree_flatten(updates)
        flat_params = None if params is None else _pytree.tree_flatten(params)[0]
        new_leaves, state = inner.update(leaves, state, params=flat_params, inplace=inplace)
        return _pytree.tree_unflatten(spec, new_leaves), state

    return GradientTransformation(init_fn, update_fn)


def scale(step_size: float) -> GradientTransformation:
    """``g * step_size`` per leaf (transform/scale.py:86-105)."""

    def update_fn(updates: Sequence, state: Any, *, params: Any = None, inplace: bool = True):  # noqa: ARG001
        if inplace:
            out = [None if g is None else g.mul_(step_size) for g in updates]
        else:
            out = [None if g is None else g.mul(step_size) for g in updates]
        return out, state

    return GradientTransformation(lambda params: (), update_fn)


def flip_sign_and_add_weight_decay(weight_decay: float = 0.0, maximize: bool = False):
    """``g <- (+/-)g + wd * p`` (alias/utils.py:52-191); identity (None) when there is nothing to do."""
    if weight_decay == 0.0 and not maximize:
        return None
    if not weight_decay >= 0.0:
        raise ValueError(f"Invalid weight_decay value: {weight_decay}")

    def update_fn(updates: Sequence, state: Any, *, params: Any = None, inplace: bool = True):
        if weight_decay != 0.0 and params is None:
            raise ValueError("params are required for weight decay")
        out = []
        for i, g in enumerate(updates):
            if g is None:
                out.append(None)
                continue
            if maximize:
                g = g.neg_() if inplace else g.neg()
            if weight_decay != 0.0:
                if not inplace:
                    g = g.add_(params[i], alpha=weight_decay) if maximize else g.add(params[i], alpha=weight_decay)
                else:
                    # a gradient outside any graph is decayed with the raw parameter values (alias/utils.py:116-121):
                    # the in-place write must not be recorded by autograd
                    p = params[i] if g.requires_grad else params[i].data
                    g = g.add_(p, alpha=weight_decay)
            out.append(g)
        return out, state

    return GradientTransformation(lambda params: (), update_fn)


# pylint: disable-next=too-many-arguments
def adam(lr: float = 1e-3, betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, *,
         eps_root: float = 0.0, moment_requires_grad: bool = False, maximize: bool = False,
         use_accelerated_op: bool = True) -> GradientTransformation:
    """Functional Adam with the accelerated op (alias/adam.py:55-137).  ``init`` / ``update`` take any pytree of
    tensors (a flat list is the common case); the state keeps flat leaf lists."""
    if not use_accelerated_op:
        raise NotImplementedError("torchopt_b200 only provides the accelerated-op path (use_accelerated_op=True)")
    if not (callable(lr) or lr >= 0.0):
        raise ValueError(f"Invalid learning rate: {lr}")
    b1, b2 = betas
    return with_flattened_tree(chain_flat(
        flip_sign_and_add_weight_decay(weight_decay, maximize),
        scale_by_accelerated_adam.flat(b1=b1, b2=b2, eps=eps, eps_root=eps_root,
                                       moment_requires_grad=moment_requires_grad),
        scale(-lr),
    ))


def add_decayed_weights(weight_decay: float = 0.0):
    """``u <- u + wd * p`` (transform/add_decayed_weights.py:195-262, no mask); identity (None) when wd == 0."""
    if not weight_decay >= 0.0:
        raise ValueError(f"Invalid weight_decay value: {weight_decay}")
    if weight_decay == 0.0:
        return None

    def update_fn(updates: Sequence, state: Any, *, params: Any = None, inplace: bool = True):
        if params is None:
            raise ValueError("params are required for weight decay")
        out = []
        for p, g in zip(params, updates):
            if g is None:
                out.append(None)
            elif inplace:
                out.append(g.add_(p if g.requires_grad else p.data, alpha=weight_decay))
            else:
                out.append(g.add(p, alpha=weight_decay))
        return out, state

    return GradientTransformation(lambda params: (), update_fn)


# pylint: disable-next=too-many-arguments
def adamw(lr: float = 1e-3, betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2, *,
          eps_root: float = 0.0, mask: Any = None, moment_requires_grad: bool = False, maximize: bool = False,
          use_accelerated_op: bool = True) -> GradientTransformation:
    """Functional AdamW with the accelerated op (alias/adamw.py:57-151); pytrees as in ``adam``."""
    if not use_accelerated_op:
        raise NotImplementedError("torchopt_b200 only provides the accelerated-op path (use_accelerated_op=True)")
    if mask is not None:
        raise NotImplementedError("masked weight decay is outside the scope of torchopt_b200")
    if not (callable(lr) or lr >= 0.0):
        raise ValueError(f"Invalid learning rate: {lr}")
    b1, b2 = betas
    return with_flattened_tree(chain_flat(
        flip_sign_and_add_weight_decay(0.0, maximize),
        scale_by_accelerated_adam.flat(b1=b1, b2=b2, eps=eps, eps_root=eps_root,
                                       moment_requires_grad=moment_requires_grad),
        add_decayed_weights(weight_decay),
        scale(-lr),
    ))


def apply_updates(params: Any, updates: Any, *, inplace: bool = True) -> Any:
    """``p + u`` per leaf (update.py:39-81): in place on ``p.data`` or as a new differentiable tensor.  ``params`` and
    ``updates`` are pytrees of the same structure (flat lists included); the result has the structure of ``params``
    (a flat sequence comes back as a list)."""
    leaves, spec = _pytree.tree_flatten(params)
    ups = _pytree.tree_flatten(updates)[0]
    if len(ups) != len(leaves):
        raise ValueError("apply_updates: params and updates have different structures")
    out = []
    for p, u in zip(leaves, ups):
        if u is None:
            out.append(p)
        elif inplace:
            p.data.add_(u)
            out.append(p)
        else:
            out.append(p.add(u))
    if spec.kind in ("list", "tuple") and all(c.kind == "leaf" for c in spec.children):
        return out
    return _pytree.tree_unflatten(spec, out)


def _fused_chain_step(step_op: AdamStep, params: list, grads: list, state: tuple) -> tuple[list, tuple]:
    """One whole-step launch for a whole adam / adamw chain + apply_updates.  ``state`` is the chain's state tuple
    (one entry per transformation, exactly one of them a ScaleByAdamState); the same structure is returned."""
    k = next(i for i, st in enumerate(state) if isinstance(st, ScaleByAdamState))
    adam_state = state[k]
    count_inc = inc_count_flat(grads, adam_state.count)
    new_params, new_mu, new_nu = step_op(params, grads, adam_state.mu, adam_state.nu, count_inc)
    return new_params, (*state[:k], ScaleByAdamState(mu=new_mu, nu=new_nu, count=count_inc), *state[k + 1:])


class _FlatGroup:
    """The leaves of one (device, parameter dtype) group: where each lives in the flat moment buffers."""

    __slots__ = ("index", "offsets", "numels", "shapes", "total", "device", "dtype", "state_dtype", "mu", "nu",
                 "cnt", "tables", "masks")

    def __init__(self, index: list, leaves: Sequence[torch.Tensor]) -> None:
        self.index = index
        self.device, self.dtype = leaves[0].device, leaves[0].dtype
        self.state_dtype = STATE_DTYPE.get(self.dtype, self.dtype)
        align = max(1, 128 // torch.empty((), dtype=self.state_dtype).element_size())
        self.numels = [p.numel() for p in leaves]
        self.shapes = [p.shape for p in leaves]
        self.offsets, total = [], 0
        for k in self.numels:
            self.offsets.append(total)
            total = (total + k + align - 1) // align * align
        self.total = total
        self.mu = self.nu = None
        self.cnt = self.tables = None      # capturable: device step counts (int64 vector) + [c1, c2, o1, c2e] tables
        self.masks = {}                    # capturable: device 0/1 vectors of the live-leaf patterns seen so far


class FlatAdamState:
    """Adam state of a fixed list of leaves with the moments of every (device, dtype) group in ONE buffer each and the
    step counts as host integers -- the state layout ``MetaAdam`` / ``FuncAdam`` / ``Adam`` use by default.

    Element for element it holds what the reference's ``ScaleByAdamState(mu, nu, count)`` holds
    (transform/scale_by_adam.py:58-63) and ``to_leaf_state`` / ``from_leaf_state`` convert without arithmetic; a step
    is one launch each way per group through ``adam_op.flat_step``: autograd sees 2L+2 inputs and L+2 outputs instead of
    4L and 3L, no per-leaf count tensors are built, and the outputs of a step come from one allocation.  Leaves that get
    no gradient in a step keep parameter, moments and count (the reference's skip semantics, adam_op.py:148-149)."""

    def __init__(self, leaves: Sequence[torch.Tensor], moment_requires_grad: bool = False, *, _empty: bool = False,
                 capturable: tuple | None = None):
        """``capturable=(b1, b2, eps_root)``: the step counts live on the DEVICE and the bias corrections come from
        host-computed device tables, so a differentiable step over this (persistent) state can be recorded in a CUDA
        graph and every replay is one real step.  The state is then carried from step to step by VALUE: each step
        reads the persistent moment buffers and writes its new moments back into them, so gradients flow through a
        step (to parameters, gradients and the moments it read) but not from one step to the next -- the contract of
        a one-step look-ahead (meta-gradient w.r.t. hyper-parameters of the current step, then ``stop_gradient``)."""
        self.capturable = capturable
        self.num = len(leaves)
        self.signature = [(p.shape, p.dtype, p.device) for p in leaves]
        by_key: dict = {}
        for i, p in enumerate(leaves):
            by_key.setdefault((p.device, p.dtype), []).append(i)
        self.groups = [_FlatGroup(idx, [leaves[i] for i in idx]) for idx in by_key.values()]
        self.counts = [0] * self.num
        self.in_capture = capturing()   # see accelerated_op.adam_op.check_capturable
        if not _empty:
            for gr in self.groups:
                gr.mu = torch.zeros(gr.total, dtype=gr.state_dtype, device=gr.device, requires_grad=moment_requires_grad)
                gr.nu = torch.zeros(gr.total, dtype=gr.state_dtype, device=gr.device, requires_grad=moment_requires_grad)
        if capturable is not None:
            b1, b2, eps_root = capturable
            for gr in self.groups:
                if gr.device.type != "cuda":
                    raise RuntimeError("Not implemented")     # no CPU path in this package
                family = 1 if gr.state_dtype == torch.float64 else 0
                gr.tables = [t.to(gr.device) for t in adam_op.bias_tables_ex(family, b1, b2, eps_root)]
                gr.cnt = torch.zeros(len(gr.index), dtype=torch.int64, device=gr.device)

    def matches(self, leaves: Sequence[torch.Tensor]) -> bool:
        return len(leaves) == self.num and all(
            p.shape == sig[0] and p.dtype == sig[1] and p.device == sig[2] for p, sig in zip(leaves, self.signature))

    # pylint: disable-next=too-many-arguments
    def step(self, hyper: tuple, params: Sequence[torch.Tensor], grads: Sequence[torch.Tensor | None], *,
             inplace: bool = False, keep_grads: bool = False) -> list:
        """One whole optimizer step (recipe + Adam + scale + apply_updates); returns the new parameter list."""
        if self.capturable is not None:
            if inplace:
                raise NotImplementedError("capturable=True with an in-place step: use Adam / FlatAdam(capturable=True)")
            return self._step_capturable(hyper, params, grads)
        if not self.in_capture and capturing():
            # bias corrections are host scalars baked into the launch: same refusal as check_capturable
            raise RuntimeError(
                "torchopt_b200: an optimizer state created OUTSIDE a CUDA-graph capture is being stepped INSIDE one; every "
                "replay would reuse the step count of the recording. Create the optimizer inside the captured function "
                "(fresh state per replay, as in a MAML inner loop) or use capturable=True (Adam / AdamW / FlatAdam(W)).")
        counts = self.counts
        after = [c + (c != INT64_MAX) if g is not None else c for c, g in zip(counts, grads)]
        b1, b2, eps, eps_root, step_size, wd_l2, wd_dec, maximize = hyper
        if len(self.groups) == 1:
            gr = self.groups[0]
            new_p, gr.mu, gr.nu = adam_op.flat_step(list(params), list(grads), gr.mu, gr.nu, gr.offsets, after, b1, b2,
                                                    eps, eps_root, step_size, wd_l2, wd_dec, maximize, inplace,
                                                    keep_grads)
        else:
            new_p = list(params)
            for gr in self.groups:
                out, gr.mu, gr.nu = adam_op.flat_step(
                    [params[i] for i in gr.index], [grads[i] for i in gr.index], gr.mu, gr.nu, gr.offsets,
                    [after[i] for i in gr.index], b1, b2, eps, eps_root, step_size, wd_l2, wd_dec, maximize, inplace,
                    keep_grads)
                for i, t in zip(gr.index, out):
                    new_p[i] = t
        self.counts = after
        return new_p

    def _step_capturable(self, hyper: tuple, params: Sequence[torch.Tensor], grads: Sequence[torch.Tensor | None]) -> list:
        b1, b2, eps, eps_root, step_size, wd_l2, wd_dec, maximize = hyper
        new_p = list(params)
        for gr in self.groups:
            g_sel = [grads[i] for i in gr.index]
            pattern = tuple(g is not None for g in g_sel)
            if not any(pattern):
                continue
            if all(pattern):
                gr.cnt.add_(1)                      # device-side increment, recorded with the graph
            else:
                mask = gr.masks.get(pattern)
                if mask is None:                    # built from fills only, so that it can appear inside a capture too
                    mask = torch.zeros(len(pattern), dtype=torch.int64, device=gr.device)
                    for k, live in enumerate(pattern):
                        if live:
                            mask[k:k + 1].fill_(1)
                    gr.masks[pattern] = mask
                gr.cnt.add_(mask)
            snap = gr.cnt.clone()                   # THIS step's counts: forward and backward read the same values
            out, mu_new, nu_new = adam_op.flat_step(
                [params[i] for i in gr.index], g_sel, gr.mu, gr.nu, gr.offsets, [0] * len(gr.index), b1, b2, eps,
                eps_root, step_size, wd_l2, wd_dec, maximize, False, False, snap, gr.tables)
            with torch.no_grad():                   # persistent state, carried by value (static addresses for replays)
                gr.mu.copy_(mu_new)
                gr.nu.copy_(nu_new)
            for i, t in zip(gr.index, out):
                new_p[i] = t
        return new_p

    def to_leaf_state(self) -> ScaleByAdamState:
        """The reference layout: per-leaf ``mu`` / ``nu`` (views of the flat buffers, differentiable) and per-leaf
        0-dim int64 device ``count`` tensors tagged with their host value."""
        mu, nu = [None] * self.num, [None] * self.num
        devices = [None] * self.num
        for gr in self.groups:
            for i, off, k, shape in zip(gr.index, gr.offsets, gr.numels, gr.shapes):
                mu[i] = gr.mu[off:off + k].view(shape)
                nu[i] = gr.nu[off:off + k].view(shape)
                devices[i] = gr.device
        if self.capturable is not None:             # the counts ARE device memory: views of the count vectors
            count = [None] * self.num
            for gr in self.groups:
                for k, i in enumerate(gr.index):
                    count[i] = gr.cnt[k]
            return ScaleByAdamState(mu=mu, nu=nu, count=count)
        return ScaleByAdamState(mu=mu, nu=nu, count=make_counts(list(self.counts), devices))

    @classmethod
    def from_leaf_state(cls, leaves: Sequence[torch.Tensor], state: ScaleByAdamState) -> "FlatAdamState":
        """Pack a reference-layout state (one differentiable ``cat`` per buffer, so gradients keep flowing to the
        tensors the caller holds); counters are read from their host mirror (one sync each if they have none)."""
        self = cls(leaves, _empty=True)
        self.counts = [count_as_int(c) for c in state.count]
        self.in_capture = all(getattr(c, "_b200_in_capture", True) for c in state.count) if self.num else capturing()
        for gr in self.groups:
            for name in ("mu", "nu"):
                src, parts, end = getattr(state, name), [], 0
                for i, off, k in zip(gr.index, gr.offsets, gr.numels):
                    if off > end:
                        parts.append(torch.zeros(off - end, dtype=gr.state_dtype, device=gr.device))
                    if src[i].dtype != gr.state_dtype or src[i].numel() != k:
                        raise ValueError("the loaded Adam state does not fit the parameters")
                    parts.append(src[i].reshape(-1))
                    end = off + k
                if gr.total > end:
                    parts.append(torch.zeros(gr.total - end, dtype=gr.state_dtype, device=gr.device))
                setattr(gr, name, torch.cat(parts) if len(parts) > 1 else parts[0].contiguous())
        return self

    def stop_gradient_(self) -> None:
        for gr in self.groups:
            if gr.mu.grad_fn is not None:
                gr.mu = gr.mu.detach().requires_grad_(True)
            if gr.nu.grad_fn is not None:
                gr.nu = gr.nu.detach().requires_grad_(True)


class _AdamStateHolder:
    """State plumbing shared by ``MetaAdam`` / ``FuncAdam`` / ``Adam``: the state lives EITHER in the flat layout
    (``FlatAdamState``, fast path) OR as the chain's state tuple in the reference layout (what ``state`` /
    ``state_dict()`` show and ``load_state_dict`` accepts).  Reading ``state`` converts flat -> reference layout without
    arithmetic; the next step then re-packs from those very tensors (one ``cat`` per buffer), so gradients w.r.t. state
    tensors the caller obtained keep flowing and in-place edits of them are honoured."""

    impl: GradientTransformation
    _flat: Any = None
    _leaf: Any = None
    _inplace_aliases = False     # in-place optimizers: the per-leaf views alias the flat buffers, both stay valid

    def _chain_state(self, adam_state: ScaleByAdamState) -> tuple:
        template = getattr(self, "_template", None)
        if template is None:
            template = self._template = self.impl.init([])
        return tuple(adam_state if isinstance(st, ScaleByAdamState) else st for st in template)

    @property
    def state(self) -> Any:
        if self._leaf is None and self._flat is not None:
            self._leaf = self._chain_state(self._flat.to_leaf_state())
        return self._leaf

    @state.setter
    def state(self, value: Any) -> None:
        self._leaf, self._flat = value, None

    def state_dict(self) -> Any:
        return self.state

    def load_state_dict(self, state: Any) -> None:
        self.state = state

    _capturable: Any = None      # (b1, b2, eps_root) for MetaAdam / FuncAdam(capturable=True)

    def _flat_state(self, leaves: list, moment_requires_grad: bool) -> FlatAdamState:
        """The flat state for ``leaves``, converting from the reference layout if that is what is current."""
        if self._capturable is not None:
            # persistent by construction: the flat buffers ARE the state; what `state` shows are views of them
            if self._flat is None:
                if self._leaf is not None:
                    raise NotImplementedError("capturable=True: load_state_dict is not supported (the state must keep "
                                              "its device addresses)")
                self._flat = FlatAdamState(leaves, moment_requires_grad, capturable=self._capturable)
            elif not self._flat.matches(leaves):
                raise ValueError("the parameters changed shape / dtype / device since the optimizer state was created")
            self._leaf = None
            return self._flat
        if self._leaf is not None and not (self._inplace_aliases and self._flat is not None):
            adam_state = next(st for st in self._leaf if isinstance(st, ScaleByAdamState))
            self._flat = FlatAdamState.from_leaf_state(leaves, adam_state)
        elif self._flat is None:
            self._flat = FlatAdamState(leaves, moment_requires_grad)
        elif not self._flat.matches(leaves):
            raise ValueError("the parameters changed shape / dtype / device since the optimizer state was created")
        self._leaf = None
        return self._flat

    def _leaf_state(self, leaves: list) -> tuple:
        """The chain's state tuple in the reference layout (A/B paths, schedules), creating it when there is none."""
        state = self.state
        if state is None:
            state = self.impl.init(leaves)
        self._leaf, self._flat = state, None
        return state

    def stop_gradient_(self) -> None:
        """``torchopt.stop_gradient(optimizer)`` (utils.py:56-92): cut the state from its graph, keep requires_grad."""
        if self._flat is not None and self._leaf is None:
            self._flat.stop_gradient_()
            return
        if self._leaf is not None:
            def cut(t: Any) -> Any:
                if isinstance(t, torch.Tensor) and t.grad_fn is not None:
                    return t.detach().requires_grad_(True)
                return t
            self.state = _pytree.tree_map(cut, self._leaf)


class CapturableCounts:
    """Device-resident step counts + bias-correction tables for an in-place optimizer whose ``step()`` is recorded in a
    CUDA graph (``capturable=True``).  Nothing about the count is a host scalar any more: ``advance`` increments the
    counts ON THE DEVICE (one tiny kernel, recorded with the graph) and the step kernel looks the corrections up in
    tables the host computed once with the reference's own expression -- so results stay bit-identical to the
    host-scalar path (tests/test_capturable_gpu.py), and every replay advances the optimizer by one real step."""

    def __init__(self, leaves: Sequence[torch.Tensor], b1: float, b2: float) -> None:
        dev, dtype = leaves[0].device, leaves[0].dtype
        if any(p.device != dev or p.dtype != dtype for p in leaves):
            raise ValueError("capturable=True needs all parameters on one device with one dtype")
        if dev.type != "cuda":
            raise RuntimeError("Not implemented")     # no CPU path in this package
        family = 1 if dtype == torch.float64 else 0    # B200_ADAM_F64 / float32 tables (fp32 and bf16 parameters)
        c1, c2 = adam_op.bias_tables(family, b1, b2)
        self.c1, self.c2 = c1.to(dev), c2.to(dev)
        self.vec = torch.zeros(len(leaves), dtype=torch.int64, device=dev)
        self.counts = list(self.vec.unbind(0))

    def advance(self, live: Sequence[int]) -> None:
        if len(live) == len(self.counts):
            self.vec.add_(1)
        else:
            for i in live:
                self.counts[i].add_(1)


def _capturable_step(step_op: AdamStep, cap: CapturableCounts, params: list, grads: list, adam_state) -> None:
    live = [i for i, g in enumerate(grads) if g is not None]
    if not live:
        return
    cap.advance(live)
    b1, b2, eps, eps_root, step_size, wd_l2, wd_dec, maximize = step_op.hyper
    adam_op.fused_step_capturable(
        [params[i].data for i in live], [grads[i] for i in live], [adam_state.mu[i] for i in live],
        [adam_state.nu[i] for i in live], [cap.counts[i] for i in live], cap.c1, cap.c2, b1, b2, eps, eps_root,
        step_size, wd_l2, wd_dec, maximize, step_op.keep_grads)


class MetaAdam(_AdamStateHolder):
    """Differentiable Adam over an ``nn.Module`` (optim/meta/adam.py:33-92 + optim/meta/base.py:35-129).

    ``step(loss)`` takes ``autograd.grad(loss, params, create_graph=True)``, runs the fused Adam update
    out of place and rebinds the module's parameters to the new (non-leaf) tensors, so an outer loss can
    back-propagate through the whole unroll.  By default the moments live in flat buffers (``FlatAdamState``): one
    launch and one autograd node per step each way; ``state`` / ``state_dict()`` still show the reference layout.
    """

    DECOUPLED = False

    # pylint: disable-next=too-many-arguments
    def __init__(self, module: nn.Module, lr: float = 1e-3, betas: tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0.0, *, eps_root: float = 0.0,
                 moment_requires_grad: bool = True, maximize: bool = False, use_accelerated_op: bool = True,
                 capturable: bool = False) -> None:
        """``capturable=True``: a differentiable step over this optimizer's PERSISTENT state may be recorded in a CUDA graph
        (device step counts + bias tables; see ``FlatAdamState``); the state is then carried between steps by value."""
        recipe = adamw if self.DECOUPLED else adam
        self.impl = recipe(lr, betas, eps, weight_decay, eps_root=eps_root, moment_requires_grad=moment_requires_grad,
                           maximize=maximize, use_accelerated_op=use_accelerated_op)
        self.fusable = not callable(lr)   # a schedule needs the chained scale_by_schedule, which is out of scope
        if capturable:
            if not self.fusable:
                raise NotImplementedError("capturable=True needs a constant learning rate")
            self._capturable = (betas[0], betas[1], eps_root)
        self.step_op = (AdamStep(betas[0], betas[1], eps, eps_root=eps_root, step_size=-lr, weight_decay=weight_decay,
                                 decoupled=self.DECOUPLED, maximize=maximize) if self.fusable else None)
        self.module = module
        self.moment_requires_grad = moment_requires_grad
        # (owner module, attribute name) for every parameter slot, in named_parameters order
        self.slots = [(m, name) for m in module.modules() for name, p in m._parameters.items() if p is not None]

    def params(self) -> list:
        return [m._parameters[name] for m, name in self.slots]

    def step(self, loss: torch.Tensor) -> None:
        flat = [m._parameters[name] for m, name in self.slots]
        if PROFILE_HOST:
            import time
            h0 = time.perf_counter()
        grads = torch.autograd.grad(loss, flat, create_graph=True, allow_unused=True)
        if PROFILE_HOST:
            h1 = time.perf_counter()
        if self.fusable and ((FUSE_STEP and FLAT_STATE) or self._capturable is not None):
            new_params = self._flat_state(flat, self.moment_requires_grad).step(self.step_op.hyper, flat, grads)
        elif self.fusable and FUSE_STEP:
            new_params, self.state = _fused_chain_step(self.step_op, flat, list(grads), self._leaf_state(flat))
        else:
            updates, self.state = self.impl.update(list(grads), self._leaf_state(flat), params=flat, inplace=False)
            new_params = apply_updates(flat, updates, inplace=False)
        for (m, name), p in zip(self.slots, new_params):
            m._parameters[name] = p
        if PROFILE_HOST:
            self.last_host_us = ((h1 - h0) * 1e6, (time.perf_counter() - h1) * 1e6)


class FuncAdam(_AdamStateHolder):
    """Functional Adam over explicit parameter lists (optim/func/base.py:36-120 + optim/func ``FuncAdam``):
    ``new_params = opt.step(loss, params)``.  ``apply_gradients`` is the same step for gradients already at hand.
    Uses the single whole-step launch (weight decay and maximize folded in) over flat moment buffers."""

    DECOUPLED = False

    # pylint: disable-next=too-many-arguments
    def __init__(self, lr: float = 1e-3, betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, *, eps_root: float = 0.0, moment_requires_grad: bool = False,
                 maximize: bool = False, use_accelerated_op: bool = True, capturable: bool = False) -> None:
        """``capturable=True``: see ``MetaAdam``."""
        recipe = adamw if self.DECOUPLED else adam
        self.impl = recipe(lr, betas, eps, weight_decay, eps_root=eps_root, moment_requires_grad=moment_requires_grad,
                           maximize=maximize, use_accelerated_op=use_accelerated_op)
        self.fusable = not callable(lr)
        if capturable:
            if not self.fusable:
                raise NotImplementedError("capturable=True needs a constant learning rate")
            self._capturable = (betas[0], betas[1], eps_root)
        self.hyper = dict(b1=betas[0], b2=betas[1], eps=eps, eps_root=eps_root, step_size=-lr if self.fusable else 0.0,
                          weight_decay=weight_decay, decoupled=self.DECOUPLED, maximize=maximize)
        self.step_hyper = AdamStep(**self.hyper).hyper
        self.moment_requires_grad = moment_requires_grad

    def apply_gradients(self, grads: Sequence, params: Sequence, inplace: bool = False) -> list:
        params = list(params)
        if self.fusable and ((FUSE_STEP and FLAT_STATE) or self._capturable is not None):
            flat = self._flat_state(params, self.moment_requires_grad)
            if inplace:   # the raw parameter values are updated; the reference writes through p.data (update.py:71-74)
                with torch.no_grad():
                    flat.step(self.step_hyper, params, list(grads), inplace=True)
                return params
            return flat.step(self.step_hyper, params, list(grads))
        if self.fusable and FUSE_STEP:
            op = AdamStep(**self.hyper, inplace=inplace)
            new_params, self.state = _fused_chain_step(op, params, list(grads), self._leaf_state(params))
            return new_params
        updates, self.state = self.impl.update(list(grads), self._leaf_state(params), params=params, inplace=inplace)
        return apply_updates(params, updates, inplace=inplace)

    def step(self, loss: torch.Tensor, params: Sequence, inplace: bool = False) -> list:
        params = list(params)
        grads = torch.autograd.grad(loss, params, create_graph=not inplace, allow_unused=True)
        return self.apply_gradients(grads, params, inplace=inplace)


class Adam(_AdamStateHolder):
    """In-place (non-differentiable) Adam over an iterable of parameters (optim/adam.py + optim/base.py:99-124)."""

    DECOUPLED = False

    # pylint: disable-next=too-many-arguments
    def __init__(self, params: Iterable[torch.Tensor], lr: float = 1e-3, betas: tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0.0, *, eps_root: float = 0.0, maximize: bool = False,
                 use_accelerated_op: bool = True, capturable: bool = False) -> None:
        """``capturable=True``: ``step()`` may be recorded in a CUDA graph together with the forward / backward of a
        training iteration (step counts live on the device, see ``CapturableCounts``); every replay is one real step."""
        self.params = list(params)
        recipe = adamw if self.DECOUPLED else adam
        self.impl = recipe(lr, betas, eps, weight_decay, eps_root=eps_root, maximize=maximize,
                           use_accelerated_op=use_accelerated_op)
        self.fusable = not callable(lr)
        self.step_op = (AdamStep(betas[0], betas[1], eps, eps_root=eps_root, step_size=-lr, weight_decay=weight_decay,
                                 decoupled=self.DECOUPLED, maximize=maximize, inplace=True) if self.fusable else None)
        self.capturable = None
        if capturable:
            if not self.fusable:
                raise NotImplementedError("capturable=True needs a constant learning rate")
            self.state = self.impl.init(self.params)
            self.capturable = CapturableCounts(self.params, betas[0], betas[1])
            self._bind_capturable_counts()
        elif self.fusable and FLAT_STATE and self.params:
            self._inplace_aliases = True
            self._flat_state(self.params, False)
        else:
            self.state = self.impl.init(self.params)

    def _adam_state_index(self) -> int:
        return next(i for i, st in enumerate(self.state) if isinstance(st, ScaleByAdamState))

    def _bind_capturable_counts(self) -> None:
        """The state's ``count`` leaves ARE the device counters (same ScaleByAdamState layout as always)."""
        k = self._adam_state_index()
        st = self.state[k]
        self.state = (*self.state[:k], ScaleByAdamState(mu=st.mu, nu=st.nu, count=self.capturable.counts),
                      *self.state[k + 1:])

    def zero_grad(self, set_to_none: bool = False) -> None:
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.detach_().zero_()

    @torch.no_grad()
    def step(self) -> None:
        grads = [p.grad for p in self.params]
        if self.capturable is not None:
            _capturable_step(self.step_op, self.capturable, self.params, grads, self.state[self._adam_state_index()])
            return
        if self.fusable and FUSE_STEP and FLAT_STATE and self.params:
            # a reference-layout view of the state that somebody read aliases the flat buffers, which are updated in
            # place; only its count tensors go stale, so it is rebuilt on the next read.  A LOADED state is re-packed.
            flat = self._flat if self._flat is not None else self._flat_state(self.params, False)
            self._leaf = None
            flat.step(self.step_op.hyper, self.params, grads, inplace=True)
        elif self.fusable and FUSE_STEP:
            _, self.state = _fused_chain_step(self.step_op, self.params, grads, self._leaf_state(self.params))
        else:
            updates, self.state = self.impl.update(grads, self._leaf_state(self.params), params=self.params, inplace=True)
            apply_updates(self.params, updates, inplace=True)


class MetaAdamW(MetaAdam):
    """Differentiable AdamW (optim/meta/adamw.py): decoupled weight decay, default 1e-2."""

    DECOUPLED = True

    def __init__(self, module: nn.Module, lr: float = 1e-3, betas: tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 1e-2, **kwargs: Any) -> None:
        super().__init__(module, lr, betas, eps, weight_decay, **kwargs)


class FuncAdamW(FuncAdam):
    """Functional AdamW (optim/func): decoupled weight decay, default 1e-2."""

    DECOUPLED = True

    def __init__(self, lr: float = 1e-3, betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, **kwargs: Any) -> None:
        super().__init__(lr, betas, eps, weight_decay, **kwargs)


class AdamW(Adam):
    """In-place AdamW (optim/adamw.py): decoupled weight decay, default 1e-2."""

    DECOUPLED = True

    def __init__(self, params: Iterable[torch.Tensor], lr: float = 1e-3, betas: tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 1e-2, **kwargs: Any) -> None:
        super().__init__(params, lr, betas, eps, weight_decay, **kwargs)
