"""``extract_state_dict`` / ``recover_state_dict`` / ``stop_gradient`` -- the three helpers the reference's MAML loops
wrap around the differentiable optimizer (examples/few-shot/maml_omniglot.py:136-165, torchopt/utils.py:56-349):

    net_state = extract_state_dict(net)            # meta-parameters before the inner loop
    opt_state = extract_state_dict(inner_opt)
    for task in tasks:
        for _ in range(K): inner_opt.step(loss(net, task))      # re-binds net's parameters to unrolled tensors
        outer_loss(net, task).backward()
        recover_state_dict(net, net_state)         # re-bind the meta-parameters: no tensor is copied or modified
        recover_state_dict(inner_opt, opt_state)   # fresh Adam state for the next task

They move references only -- no arithmetic, nothing on the GPU unless a copy mode is asked for -- and work with
``MetaAdam`` / ``MetaAdamW`` (a tuple of per-transformation states) and with ``FlatMetaAdam`` (three references: the
flat parameter buffer and the flat moments).  Visualisation annotations of the reference are
out of scope.
"""
from __future__ import annotations

from typing import Any, NamedTuple

import torch
from torch import nn

from torchopt_b200 import _pytree

__all__ = ["ModuleState", "extract_state_dict", "recover_state_dict", "stop_gradient"]


class ModuleState(NamedTuple):
    """Parameters and buffers of a module, one ``{name: tensor}`` dict per sub-module that owns any
    (utils.py:46-52)."""

    params: tuple
    buffers: tuple
    detach_buffers: bool = False


def _is_meta_optimizer(obj: Any) -> bool:
    return callable(getattr(obj, "state_dict", None)) and callable(getattr(obj, "load_state_dict", None)) \
        and callable(getattr(obj, "step", None)) and not isinstance(obj, nn.Module)


def _containers(module: nn.Module, with_buffers: bool = True) -> tuple[list, list]:
    """The live ``_parameters`` / ``_buffers`` dicts of every distinct sub-module that has entries, in ``modules()``
    order (the same walk ``extract`` and ``recover`` use, so the two line up)."""
    params, buffers = [], []
    for sub in module.modules():   # modules() already yields each module once
        if len(sub._parameters) > 0:
            params.append(sub._parameters)
        if with_buffers and len(sub._buffers) > 0:
            buffers.append(sub._buffers)
    return params, buffers


def _replicator(by: str, device: Any):
    by = {"ref": "reference", "clone": "copy", "deepclone": "deepcopy"}.get(by, by)
    if by not in ("reference", "copy", "deepcopy"):
        raise ValueError(f"unknown copy mode: {by!r}")
    dev = None if device is None else torch.device(device)

    def move(t: torch.Tensor) -> torch.Tensor:
        return t if dev is None else t.to(device=dev)

    def deep(t: torch.Tensor) -> torch.Tensor:
        out = move(t.clone()).detach_()
        if isinstance(t, nn.Parameter):
            return nn.Parameter(out, requires_grad=t.requires_grad)
        return out.requires_grad_(t.requires_grad)

    if by == "reference":
        return move, deep
    if by == "copy":
        return (lambda t: move(t.clone())), deep
    return deep, deep


# pylint: disable-next=too-many-arguments
def extract_state_dict(target: Any, *, by: str = "reference", device: Any = None, with_buffers: bool = True,
                       detach_buffers: bool = False) -> Any:
    """State of a module (``ModuleState``) or of a differentiable optimizer (whatever its ``state_dict()`` holds, with
    every tensor replicated by ``by``): ``'reference'`` (default; the same tensors), ``'copy'`` (clones that stay in
    the autograd graph) or ``'deepcopy'`` (detached clones)."""
    replicate, deep = _replicator(by, device)
    if isinstance(target, nn.Module):
        params, buffers = _containers(target, with_buffers)
        buf_fn = deep if detach_buffers else replicate
        return ModuleState(
            params=tuple(type(c)((k, replicate(v)) for k, v in c.items() if isinstance(v, torch.Tensor))
                         for c in params),
            buffers=tuple(type(c)((k, buf_fn(v)) for k, v in c.items() if isinstance(v, torch.Tensor))
                          for c in buffers),
            detach_buffers=detach_buffers)
    if _is_meta_optimizer(target):
        return _pytree.tree_map(lambda t: replicate(t) if isinstance(t, torch.Tensor) else t, target.state_dict())
    raise TypeError(f"Unexpected class of {target}")


def recover_state_dict(target: Any, state: Any) -> None:
    """Put an extracted state back: a module's parameter / buffer slots are re-bound to the saved tensors (nothing is
    copied into existing tensors); an optimizer gets ``load_state_dict(state)``."""
    if isinstance(target, nn.Module):
        if not isinstance(state, ModuleState):
            raise TypeError("recover_state_dict(module, state): state must come from extract_state_dict(module)")
        params, buffers = _containers(target, with_buffers=True)
        saved_buffers = state.buffers
        if state.detach_buffers:
            _, deep = _replicator("deepcopy", None)
            saved_buffers = tuple(type(c)((k, deep(v)) for k, v in c.items()) for c in saved_buffers)
        for live, saved in list(zip(params, state.params)) + list(zip(buffers, saved_buffers)):
            live.update(saved)
    elif _is_meta_optimizer(target):
        target.load_state_dict(state)
    else:
        raise TypeError(f"Unexpected class of {target}")


def stop_gradient(target: Any) -> None:
    """Detach every tensor of a module / optimizer state / ``ModuleState`` / pytree from its graph, IN PLACE, keeping
    ``requires_grad`` (utils.py:56-92): the truncation point of a truncated unroll."""
    custom = getattr(target, "stop_gradient_", None)
    if callable(custom):          # layouts that hold views of one buffer cut the buffer and re-derive the views
        custom()
        return
    if isinstance(target, ModuleState):
        tree: Any = (target.params, target.buffers)
    elif isinstance(target, nn.Module):
        tree = tuple(p for sub in target.modules() for p in sub._parameters.values() if p is not None)
    elif _is_meta_optimizer(target):
        tree = target.state_dict()
    else:
        tree = target

    def cut(t: Any) -> Any:
        if isinstance(t, torch.Tensor) and t.grad_fn is not None:   # graph leaves have nothing to cut
            t.detach_().requires_grad_(True)
        return t

    _pytree.tree_map(cut, tree)
