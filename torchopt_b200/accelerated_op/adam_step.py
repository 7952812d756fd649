"""``AdamStep`` -- the whole differentiable optimizer step as ONE autograd node and ONE launch each way.

Fuses what the reference runs as three chained transformations plus ``apply_updates`` for every leaf
(/root/reference/torchopt/alias/adam.py:116-137 -> transform/scale_by_adam.py:304-348 (3 launches/leaf) ->
transform/scale.py:86-105 (``g.mul(-lr)``, 1 launch/leaf) -> update.py:78-79 (``p.add(u)``, 1 launch/leaf)):

    forward :  (p, g, mu, nu) for all leaves -> (p', mu', nu')      28 B/elem fp32, 1 launch
    backward:  (dp', dmu', dnu')             -> (dp = dp', dg, dmu, dnu)  36 B/elem fp32, 1 launch

Results are bit-identical to running ``scale_by_accelerated_adam`` + ``scale(step_size)`` + ``apply_updates`` with
torch ops: the kernel performs the same rounded fp32/fp64 operations in the same order (``us = u * S``, ``p' = p + us``;
backward ``du = dp' * S``), and the two gradient accumulations it absorbs have two addends each.
SURVEY.md section 8f row n1.  Weight decay / ``maximize`` are not folded in: callers fall back to the chained path.
"""
from __future__ import annotations

from typing import Any, Sequence

import torch

from torchopt_b200._native import adam_op
from torchopt_b200.accelerated_op.adam_op import count_as_int


class AdamStep:  # pylint: disable=too-few-public-methods
    """``AdamStep(b1, b2, eps, eps_root, step_size)(params, grads, mu, nu, counts) -> (new_params, new_mu, new_nu)``."""

    class Op(torch.autograd.Function):  # pylint: disable=abstract-method
        """Tensor arguments ``p_0.., g_0.., mu_0.., nu_0..``; outputs ``p'_0.., mu'_0.., nu'_0..``."""

        @staticmethod
        def forward(ctx: Any, hyper: tuple, counts: tuple, *tensors: torch.Tensor) -> Any:
            b1, b2, eps, eps_root, step_size = hyper
            num = len(counts)
            params, grads = tensors[:num], tensors[num:2 * num]
            mus, nus = tensors[2 * num:3 * num], tensors[3 * num:]
            new_p, new_mu, new_nu, _ = adam_op.fused_step(
                list(params), list(grads), list(mus), list(nus), b1, b2, eps, eps_root, step_size, list(counts),
                False, False)
            ctx.hyper, ctx.counts = hyper, counts
            ctx.set_materialize_grads(False)
            ctx.save_for_backward(*grads, *new_mu, *new_nu)   # u is recomputed in the backward kernel
            return (*new_p, *new_mu, *new_nu)

        @staticmethod
        def backward(ctx: Any, *douts: torch.Tensor | None) -> Any:
            b1, b2, eps, eps_root, step_size = ctx.hyper
            num = len(ctx.counts)
            saved = ctx.saved_tensors
            grads, new_mu, new_nu = saved[:num], saved[num:2 * num], saved[2 * num:]
            dparams, dmu_ext, dnu_ext = douts[:num], douts[num:2 * num], douts[2 * num:]
            need = ctx.needs_input_grad[2:]
            need_p, need_g = need[:num], need[num:2 * num]
            need_mu, need_nu = need[2 * num:3 * num], need[3 * num:]
            dg, dmu, dnu = adam_op.fused_step_backward(
                list(dparams), [None] * num, list(dmu_ext), list(dnu_ext), list(new_mu), list(new_nu), list(grads),
                list(need_g), list(need_mu), list(need_nu), b1, b2, eps, eps_root, step_size, list(ctx.counts))
            # p' = p + us  =>  dL/dp = dL/dp' : hand the incoming tensor straight back, no bytes move
            dp = [d if n else None for d, n in zip(dparams, need_p)]
            return (None, None, *dp, *dg, *dmu, *dnu)

    # pylint: disable-next=too-many-arguments
    def __init__(self, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8, *, eps_root: float = 0.0,
                 step_size: float = -1e-3, inplace: bool = False) -> None:
        self.hyper = (b1, b2, eps, eps_root, step_size)
        self.inplace = inplace

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
            b1, b2, eps, eps_root, step_size = self.hyper
            adam_op.fused_step([params[i].data for i in live], [grads[i] for i in live], [mus[i] for i in live],
                               [nus[i] for i in live], b1, b2, eps, eps_root, step_size, list(host_counts), True, False)
            return new_p, new_mu, new_nu
        k = len(live)
        flat = self.Op.apply(self.hyper, host_counts, *[params[i] for i in live], *[grads[i] for i in live],
                             *[mus[i] for i in live], *[nus[i] for i in live])
        for j, i in enumerate(live):
            new_p[i], new_mu[i], new_nu[i] = flat[j], flat[k + j], flat[2 * k + j]
        return new_p, new_mu, new_nu
