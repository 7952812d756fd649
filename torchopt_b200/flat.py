"""Flat-buffer layout for parameters, gradients and Adam moments.

The reference keeps a pytree of independent leaves everywhere (optim/meta/base.py:58-96, optim/base.py:99-124), so every
optimizer step is O(leaves) Python work, O(leaves) autograd inputs/outputs and a multi-tensor leaf table.  With 180 GB
of HBM per GPU there is no reason not to lay a model's parameters, their gradients and both moments out as FOUR
contiguous buffers:

* the whole optimizer step becomes a single-range launch (the 1-leaf table: no leaf search, one ragged tail instead of
  one per leaf) and ONE autograd node with 4 tensor inputs / 3 outputs instead of 4L / 3L;
* the gradient of a loss w.r.t. the flat parameter buffer arrives as one flat tensor (autograd concatenates the view
  gradients once, 8 B/elem), which is also the meta-gradient all-reduce bucket -- no flatten, no copy-back;
* saving / restoring optimizer state per task (``torchopt.recover_state_dict``, utils.py:309-349) swaps three tensor
  references instead of walking a tree.

``FlatLayout`` places each leaf at a 128-byte aligned offset.  ``FlatMetaAdam`` is the differentiable optimizer over an
``nn.Module`` whose parameter slots are re-bound to views of the current flat buffer; ``FlatAdam`` is the in-place
optimizer whose ``p.data`` / ``p.grad`` are views of flat buffers.  Element for element both compute exactly what
``MetaAdam`` / ``Adam`` compute (same kernels, same order), with one documented difference: a leaf that receives no
gradient counts as a zero gradient (its moments decay) instead of being skipped, because a flat gradient has no holes.
"""
from __future__ import annotations

from typing import Any, Iterable, Sequence

import torch
from torch import nn

from torchopt_b200.accelerated_op.adam_step import AdamStep
from torchopt_b200.optim import CapturableCounts, _capturable_step, _fused_chain_step, adam, adamw
from torchopt_b200.transform import ScaleByAdamState

__all__ = ["FlatLayout", "FlatMetaAdam", "FlatMetaAdamW", "FlatAdam", "FlatAdamW"]


class FlatLayout:
    """Offsets of ``leaves`` inside one contiguous buffer of their common dtype on their common device.

    Every leaf starts on an ``align_bytes`` boundary (default 128 B: a full L2 sector pair, and enough for the 256-bit
    vector path and for cuDNN / cuBLAS alignment requirements); the gaps are padding elements that always hold 0.
    """

    def __init__(self, leaves: Sequence[torch.Tensor], align_bytes: int = 128) -> None:
        leaves = list(leaves)
        if not leaves:
            raise ValueError("FlatLayout needs at least one leaf")
        self.dtype, self.device = leaves[0].dtype, leaves[0].device
        for p in leaves:
            if p.dtype != self.dtype or p.device != self.device:
                raise ValueError("FlatLayout: all leaves must share one dtype and one device")
        align = max(1, align_bytes // leaves[0].element_size())
        self.shapes = [tuple(p.shape) for p in leaves]
        self.numels = [p.numel() for p in leaves]
        self.offsets: list[int] = []
        self.chunks: list[int] = []      # split sizes: leaf, pad, leaf, pad, ... (zero-length pads omitted)
        self.keep: list[int] = []        # index into `chunks` of every leaf
        total = 0
        for k in self.numels:
            self.offsets.append(total)
            self.keep.append(len(self.chunks))
            self.chunks.append(k)
            end = total + k
            total = (end + align - 1) // align * align
            if total > end:
                self.chunks.append(total - end)
        self.total = total

    def pack(self, leaves: Sequence[torch.Tensor]) -> torch.Tensor:
        """One flat tensor holding ``leaves`` at their offsets (one ``cat`` kernel; differentiable w.r.t. the leaves,
        the backward is a set of views)."""
        parts, leaf_at = [], dict(zip(self.keep, leaves))
        for j, k in enumerate(self.chunks):
            if j in leaf_at:
                if leaf_at[j].numel() != k:
                    raise ValueError("FlatLayout.pack: leaf sizes differ from the layout")
                parts.append(leaf_at[j].reshape(-1))
            else:
                parts.append(torch.zeros(k, dtype=self.dtype, device=self.device))
        return torch.cat(parts)

    def views(self, flat: torch.Tensor) -> list[torch.Tensor]:
        """The leaves as views of ``flat`` (no copy).  Differentiable: the gradient w.r.t. ``flat`` is assembled by
        ONE concatenation of the view gradients (zeros for views that got none and for the padding)."""
        if flat.numel() != self.total:
            raise ValueError(f"flat buffer has {flat.numel()} elements, layout needs {self.total}")
        pieces = flat.split_with_sizes(self.chunks)
        return [pieces[j].view(shape) for j, shape in zip(self.keep, self.shapes)]

    def raw_views(self, flat: torch.Tensor) -> list[torch.Tensor]:
        """Views outside autograd (for ``.data`` / ``.grad`` re-pointing)."""
        flat = flat.detach()
        return [flat[o:o + k].view(shape) for o, k, shape in zip(self.offsets, self.numels, self.shapes)]


def _check_padding_safe(eps: float, eps_root: float) -> None:
    """``FlatMetaAdam`` / ``FlatAdam`` launch ONE range over the whole buffer, padding included.  Padding holds
    g = mu = nu = 0, for which the update is 0 / (sqrt(0 + eps_root) + eps): exactly 0 unless both epsilons are 0, when it
    is NaN -- and a NaN in the padding of the flat parameters would poison every whole-buffer reduction (norms, clipping,
    an all-reduce sanity check).  (The default optimizers -- ``MetaAdam`` / ``Adam`` over ``FlatAdamState`` -- launch per
    leaf range and never touch padding.)"""
    if eps == 0.0 and eps_root == 0.0:
        raise ValueError("FlatMetaAdam / FlatAdam need eps > 0 or eps_root > 0: with both zero the update of the zero "
                         "padding between leaves is 0/0 = NaN (use MetaAdam / Adam, which never touch padding)")


def _module_slots(module: nn.Module) -> list:
    return [(m, name) for m in module.modules() for name, p in m._parameters.items() if p is not None]


class FlatMetaAdam:
    """Differentiable Adam over an ``nn.Module`` with ALL parameters in one flat buffer (optim/meta/adam.py +
    optim/meta/base.py:35-129 re-laid-out).

        inner = FlatMetaAdam(net, lr=1e-3)          # packs the parameters, re-binds net's slots to views
        for _ in range(K):
            inner.step(loss_fn(net(x), y))          # 1 autograd.grad w.r.t. ONE tensor + ONE launch
        outer_loss(net).backward()                  # reaches the original (meta-)parameters through the pack
        inner.restore()                             # give net its original parameter tensors back

    Per step: ``autograd.grad(loss, flat_p, create_graph=True)`` (one flat gradient), one whole-step launch over
    ``(flat_p, flat_g, flat_mu, flat_nu)``, new views.  Per unrolled step the backward is one launch as well.
    """

    DECOUPLED = False

    # pylint: disable-next=too-many-arguments
    def __init__(self, module: nn.Module, lr: float = 1e-3, betas: tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0.0, *, eps_root: float = 0.0,
                 moment_requires_grad: bool = True, maximize: bool = False, use_accelerated_op: bool = True) -> None:
        if callable(lr):
            raise NotImplementedError("learning-rate schedules are outside the scope of torchopt_b200")
        _check_padding_safe(eps, eps_root)
        recipe = adamw if self.DECOUPLED else adam
        self.impl = recipe(lr, betas, eps, weight_decay, eps_root=eps_root, moment_requires_grad=moment_requires_grad,
                           maximize=maximize, use_accelerated_op=use_accelerated_op)
        self.step_op = AdamStep(betas[0], betas[1], eps, eps_root=eps_root, step_size=-lr, weight_decay=weight_decay,
                                decoupled=self.DECOUPLED, maximize=maximize)
        self.module = module
        self.slots = _module_slots(module)
        self.original = [m._parameters[name] for m, name in self.slots]
        self.layout = FlatLayout(self.original)
        self.flat_p = self.layout.pack(self.original)
        self.state: Any = None
        self._bind(self.flat_p)

    def _bind(self, flat: torch.Tensor) -> None:
        for (m, name), v in zip(self.slots, self.layout.views(flat)):
            m._parameters[name] = v

    def restore(self) -> None:
        """Re-bind the module's slots to the tensors it had when this optimizer was built."""
        for (m, name), p in zip(self.slots, self.original):
            m._parameters[name] = p

    def params(self) -> list:
        return [m._parameters[name] for m, name in self.slots]

    def step(self, loss: torch.Tensor) -> None:
        if self.state is None:
            self.state = self.impl.init([self.flat_p])
        (grad,) = torch.autograd.grad(loss, [self.flat_p], create_graph=True, allow_unused=True)
        if grad is None:
            return
        (new_flat,), self.state = _fused_chain_step(self.step_op, [self.flat_p], [grad], self.state)
        self.flat_p = new_flat
        self._bind(new_flat)

    def stop_gradient_(self) -> None:
        """Cut the unroll here (torchopt.stop_gradient semantics): the flat parameters and the moments become graph
        leaves that keep ``requires_grad``, and the module's slots are re-bound to views of the detached buffer."""
        self.flat_p = self.flat_p.detach().requires_grad_(True)
        if self.state is not None:
            def cut(t: Any) -> Any:
                if isinstance(t, torch.Tensor) and t.grad_fn is not None:
                    return t.detach().requires_grad_(True)
                return t
            self.state = tuple(
                ScaleByAdamState(mu=[cut(t) for t in st.mu], nu=[cut(t) for t in st.nu], count=st.count)
                if isinstance(st, ScaleByAdamState) else st for st in self.state)
        self._bind(self.flat_p)

    # per-task state save / restore is a reference swap (utils.py:309-349 walks the tree and clones)
    def state_dict(self) -> Any:
        return (self.flat_p, self.state)

    def load_state_dict(self, state: Any) -> None:
        self.flat_p, self.state = state
        self._bind(self.flat_p)


class FlatMetaAdamW(FlatMetaAdam):
    """Differentiable AdamW over one flat buffer: decoupled weight decay, default 1e-2 (optim/meta/adamw.py)."""

    DECOUPLED = True

    def __init__(self, module: nn.Module, lr: float = 1e-3, betas: tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 1e-2, **kwargs: Any) -> None:
        super().__init__(module, lr, betas, eps, weight_decay, **kwargs)


class FlatAdam:
    """In-place Adam whose parameters, gradients and moments live in four flat buffers (optim/adam.py +
    optim/base.py:99-124 re-laid-out).  ``p.data`` and ``p.grad`` of every parameter become views of ``flat_p`` /
    ``flat_g`` (values preserved), so ``loss.backward()`` accumulates straight into ``flat_g`` -- which is also the
    all-reduce bucket for data-parallel training -- and ``step()`` is ONE single-range launch, 28 B/elem (the gradients
    are left as they are; the reference's in-place chain overwrites them with the updates).

    If some ``p.grad`` has been replaced (e.g. ``zero_grad(set_to_none=True)`` by foreign code followed by a fresh
    backward), ``step()`` first re-attaches the views and copies such gradients in; missing ones count as zero.
    """

    DECOUPLED = False

    # pylint: disable-next=too-many-arguments
    def __init__(self, params: Iterable[torch.Tensor], lr: float = 1e-3, betas: tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0.0, *, eps_root: float = 0.0, maximize: bool = False,
                 use_accelerated_op: bool = True, capturable: bool = False) -> None:
        """``capturable=True``: the step count lives on the device, so ``step()`` can be recorded in a CUDA graph with
        the rest of a training iteration (``zero_grad`` + forward + backward + ``step``); each replay is a real step."""
        if callable(lr):
            raise NotImplementedError("learning-rate schedules are outside the scope of torchopt_b200")
        _check_padding_safe(eps, eps_root)
        self.params = list(params)
        recipe = adamw if self.DECOUPLED else adam
        self.impl = recipe(lr, betas, eps, weight_decay, eps_root=eps_root, maximize=maximize,
                           use_accelerated_op=use_accelerated_op)
        # the flat gradient buffer belongs to this optimizer (zero_grad clears it), so the reference's side effect of
        # turning gradients into updates is not reproduced: the step writes p, mu, nu only -- 28 B/elem
        self.step_op = AdamStep(betas[0], betas[1], eps, eps_root=eps_root, step_size=-lr, weight_decay=weight_decay,
                                decoupled=self.DECOUPLED, maximize=maximize, inplace=True, keep_grads=True)
        self.layout = FlatLayout(self.params)
        with torch.no_grad():
            self.flat_p = self.layout.pack([p.detach() for p in self.params])
        self.flat_g = torch.zeros_like(self.flat_p)
        self.p_views = self.layout.raw_views(self.flat_p)
        self.g_views = self.layout.raw_views(self.flat_g)
        for p, pv, gv in zip(self.params, self.p_views, self.g_views):
            if p.grad is not None:
                gv.copy_(p.grad)
            p.data = pv
            p.grad = gv
        self.state = self.impl.init([self.flat_p])
        self.capturable = None
        if capturable:
            self.capturable = CapturableCounts([self.flat_p], betas[0], betas[1])
            k = next(i for i, st in enumerate(self.state) if isinstance(st, ScaleByAdamState))
            st = self.state[k]
            self.state = (*self.state[:k], ScaleByAdamState(mu=st.mu, nu=st.nu, count=self.capturable.counts),
                          *self.state[k + 1:])

    def zero_grad(self, set_to_none: bool = False) -> None:  # noqa: ARG002  the flat buffer is never dropped
        """One memset of the flat gradient buffer; the ``.grad`` views stay attached."""
        self.flat_g.zero_()
        for p, gv in zip(self.params, self.g_views):
            p.grad = gv

    def _gather_foreign_grads(self) -> None:
        for p, gv in zip(self.params, self.g_views):
            if p.grad is gv:
                continue
            if p.grad is None:
                gv.zero_()
            elif p.grad.data_ptr() != gv.data_ptr():
                gv.copy_(p.grad)
            p.grad = gv

    @torch.no_grad()
    def step(self) -> None:
        self._gather_foreign_grads()
        if self.capturable is not None:
            st = next(s for s in self.state if isinstance(s, ScaleByAdamState))
            _capturable_step(self.step_op, self.capturable, [self.flat_p], [self.flat_g], st)
            return
        _, self.state = _fused_chain_step(self.step_op, [self.flat_p], [self.flat_g], self.state)

    def state_dict(self) -> Any:
        """The chain's state tuple (one entry per transformation; the ``ScaleByAdamState`` holds the flat moments)."""
        return self.state

    def load_state_dict(self, state: Any) -> None:
        self.state = state


class FlatAdamW(FlatAdam):
    """In-place AdamW over flat buffers: decoupled weight decay, default 1e-2 (optim/adamw.py)."""

    DECOUPLED = True

    def __init__(self, params: Iterable[torch.Tensor], lr: float = 1e-3, betas: tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 1e-2, **kwargs: Any) -> None:
        super().__init__(params, lr, betas, eps, weight_decay, **kwargs)
