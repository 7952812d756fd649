"""``AdamStep`` -- the whole differentiable optimizer step as ONE autograd node and ONE launch each way.

Fuses what the reference runs as three chained transformations plus ``apply_updates`` for every leaf
(/root/reference/torchopt/alias/adam.py:116-137 -> transform/scale_by_adam.py:304-348 (3 launches/leaf) ->
transform/scale.py:86-105 (``g.mul(-lr)``, 1 launch/leaf) -> update.py:78-79 (``p.add(u)``, 1 launch/leaf)):

    forward :  (p, g, mu, nu) for all leaves -> (p', mu', nu')      28 B/elem fp32, 1 launch
    backward:  (dp', dmu', dnu')             -> (dp = dp', dg, dmu, dnu)  36 B/elem fp32, 1 launch

Results are bit-identical to running ``scale_by_accelerated_adam`` + ``scale(step_size)`` + ``apply_updates`` with
torch ops: the kernel performs the same rounded fp32/fp64 operations in the same order (``us = u * S``, ``p' = p + us``;
backward ``du = dp' * S``), and the two gradient accumulations it absorbs have two addends each.
SURVEY.md section 8f rows n1 and n3: L2 weight decay (``adam``), decoupled weight decay (``adamw``) and ``maximize``
are folded into the same launch (``x.add(y, alpha=a)`` evaluated as ATen does on CUDA: fma(a, y, x)).
"""
from __future__ import annotations

import os
from typing import Any, Sequence

import torch

from torchopt_b200._native import adam_op
from torchopt_b200.accelerated_op.adam_op import count_as_int

# A/B switch for parity tests: True = native autograd node (torch_shim.cpp:StepNode), False = the Python
# autograd.Function below (AdamStep.Op), kept as the readable specification of the node.
NATIVE_NODE = os.environ.get("B200_ADAM_PY_NODES", "") in ("", "0")


class AdamStep:  # pylint: disable=too-few-public-methods
    """``AdamStep(...)(params, grads, mu, nu, counts) -> (new_params, new_mu, new_nu)`` for a whole recipe:
    ``adam`` (optional L2 ``weight_decay``), ``adamw`` (``decoupled=True``), optional ``maximize``."""

    class Op(torch.autograd.Function):  # pylint: disable=abstract-method
        """Tensor arguments ``p_0.., g_0.., mu_0.., nu_0..``; outputs ``p'_0.., mu'_0.., nu'_0..``."""

        @staticmethod
        def forward(ctx: Any, hyper: tuple, counts: tuple, *tensors: torch.Tensor) -> Any:
            b1, b2, eps, eps_root, step_size, wd_l2, wd_dec, maximize = hyper
            num = len(counts)
            params, grads = tensors[:num], tensors[num:2 * num]
            mus, nus = tensors[2 * num:3 * num], tensors[3 * num:]
            new_p, new_mu, new_nu, _ = adam_op.fused_step(
                list(params), list(grads), list(mus), list(nus), b1, b2, eps, eps_root, step_size, list(counts),
                False, False, wd_l2, wd_dec, maximize)
            ctx.hyper, ctx.counts = hyper, counts
            ctx.set_materialize_grads(False)
            # u is recomputed in the backward kernel; the raw parameters are needed only to re-derive g + wd*p
            extra = params if wd_l2 != 0.0 else ()
            ctx.save_for_backward(*grads, *new_mu, *new_nu, *extra)
            return (*new_p, *new_mu, *new_nu)

        @staticmethod
        def backward(ctx: Any, *douts: torch.Tensor | None) -> Any:
            b1, b2, eps, eps_root, step_size, wd_l2, wd_dec, maximize = ctx.hyper
            num = len(ctx.counts)
            saved = ctx.saved_tensors
            grads, new_mu, new_nu = saved[:num], saved[num:2 * num], saved[2 * num:3 * num]
            params = list(saved[3 * num:]) if wd_l2 != 0.0 else [None] * num
            dparams, dmu_ext, dnu_ext = douts[:num], douts[num:2 * num], douts[2 * num:]
            need = ctx.needs_input_grad[2:]
            need_p, need_g = need[:num], need[num:2 * num]
            need_mu, need_nu = need[2 * num:3 * num], need[3 * num:]
            dg, dmu, dnu, dp = adam_op.fused_step_backward(
                list(dparams), [None] * num, list(dmu_ext), list(dnu_ext), list(new_mu), list(new_nu), list(grads),
                list(need_g), list(need_mu), list(need_nu), b1, b2, eps, eps_root, step_size, list(ctx.counts),
                params, list(need_p), wd_l2, wd_dec, maximize)
            return (None, None, *dp, *dg, *dmu, *dnu)

    # pylint: disable-next=too-many-arguments
    def __init__(self, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8, *, eps_root: float = 0.0,
                 step_size: float = -1e-3, weight_decay: float = 0.0, decoupled: bool = False, maximize: bool = False,
                 inplace: bool = False, keep_grads: bool = False) -> None:
        """``inplace=True`` is the non-differentiable ``Optimizer.step`` form: parameters and moments are updated in
        place and -- as in the reference, whose in-place chain turns each gradient tensor into the scaled update --
        the gradients are overwritten, unless ``keep_grads=True`` (skips that 4 B/elem store: 28 instead of 32)."""
        wd_l2, wd_dec = (0.0, weight_decay) if decoupled else (weight_decay, 0.0)
        self.hyper = (b1, b2, eps, eps_root, step_size, wd_l2, wd_dec, bool(maximize))
        self.inplace = inplace
        self.keep_grads = keep_grads

    def __call__(self, params: Sequence[torch.Tensor], grads: Sequence[torch.Tensor | None],
                 mus: Sequence[torch.Tensor], nus: Sequence[torch.Tensor], counts: Sequence[Any],
                 ) -> tuple[list, list, list]:
        """Leaves whose gradient is None keep their parameter and moments (reference: adam_op.py:148-149,
        update.py:74-79)."""
        num = len(params)
        live = [i for i in range(num) if grads[i] is not None]
        new_p, new_mu, new_nu = list(params), list(mus), list(nus)
        if not live:
            return new_p, new_mu, new_nu
        host_counts = tuple(count_as_int(counts[i]) for i in live)
        if self.inplace:
            b1, b2, eps, eps_root, step_size, wd_l2, wd_dec, maximize = self.hyper
            adam_op.fused_step([params[i].data for i in live], [grads[i] for i in live], [mus[i] for i in live],
                               [nus[i] for i in live], b1, b2, eps, eps_root, step_size, list(host_counts), True, False,
                               wd_l2, wd_dec, maximize, self.keep_grads)
            return new_p, new_mu, new_nu
        k = len(live)
        if NATIVE_NODE:
            b1, b2, eps, eps_root, step_size, wd_l2, wd_dec, maximize = self.hyper
            flat = adam_op.step_autograd([params[i] for i in live], [grads[i] for i in live], [mus[i] for i in live],
                                         [nus[i] for i in live], b1, b2, eps, eps_root, step_size, list(host_counts),
                                         wd_l2, wd_dec, maximize)
        else:
            flat = self.Op.apply(self.hyper, host_counts, *[params[i] for i in live], *[grads[i] for i in live],
                                 *[mus[i] for i in live], *[nus[i] for i in live])
        for j, i in enumerate(live):
            new_p[i], new_mu[i], new_nu[i] = flat[j], flat[k + j], flat[2 * k + j]
        return new_p, new_mu, new_nu
