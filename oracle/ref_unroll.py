"""TEST INFRASTRUCTURE -- never imported by the product (torchopt_b200/).

How the reference runs a differentiable Adam step on CPU, for BASELINE config [0] ("torchopt.adam(use_accelerated_op=
True) fwd+bwd on a synthetic 1M-param MLP, 5-step unrolled meta-grad, on CPU (adam_op_cpu)"): the reference's COMPILED
CPU kernels (oracle/_ref/_C.so, unmodified sources, built by oracle/build_ref.py) driven by a restatement of its
Python glue --

    AdamOp.MuOp / NuOp / UpdatesOp   torchopt/accelerated_op/adam_op.py:36-121   three autograd Functions per leaf
    AdamOp.__call__                  adam_op.py:140-181                          mu', nu', u for one leaf
    scale_by_accelerated_adam        transform/scale_by_adam.py:304-348          Python loop over the leaves, inc_count
    scale(-lr), apply_updates        transform/scale.py:86-105, update.py:69-81  one torch op per leaf each
    MetaOptimizer.step               optim/meta/base.py:58-96                    autograd.grad(create_graph=True), re-bind

-- because `import torchopt` itself needs `optree`, which this image does not have.  The
restatement is pinned: tests/test_oracle.py::test_ref_unroll_reproduces_golden reproduces tests/golden/adam_unroll_*.npz,
which were generated with the reference's own adam_op.py imported by file path.

Used by bench.py's CPU-baseline leg (`extra.config0_mlp1m`) and by tests only.
"""
from __future__ import annotations

from typing import Any, Sequence

import torch

from oracle import adam_oracle as ao


def make_ref_adam_op(ref_ops: Any):
    """The reference's AdamOp over ``ref_ops`` (= oracle/_ref ``_C.adam_op``)."""

    class MuOp(torch.autograd.Function):  # adam_op.py:36-60
        @staticmethod
        def forward(ctx, updates, mu, b1):
            ctx.b1 = b1
            ctx.save_for_backward(updates, mu)
            return ref_ops.forward_mu(updates, mu, b1)

        @staticmethod
        def backward(ctx, dmu):
            updates, mu = ctx.saved_tensors
            out = ref_ops.backward_mu(dmu, updates, mu, ctx.b1)
            return out[0], out[1], None

    class NuOp(torch.autograd.Function):  # adam_op.py:62-86
        @staticmethod
        def forward(ctx, updates, nu, b2):
            ctx.b2 = b2
            ctx.save_for_backward(updates, nu)
            return ref_ops.forward_nu(updates, nu, b2)

        @staticmethod
        def backward(ctx, dnu):
            updates, nu = ctx.saved_tensors
            out = ref_ops.backward_nu(dnu, updates, nu, ctx.b2)
            return out[0], out[1], None

    class UpdatesOp(torch.autograd.Function):  # adam_op.py:88-121
        @staticmethod
        def forward(ctx, new_mu, new_nu, others):
            b1, b2, eps, eps_root, count = others
            new_updates = ref_ops.forward_updates(new_mu, new_nu, b1, b2, eps, eps_root, count)
            ctx.others = (b1, b2, eps_root, count)
            ctx.save_for_backward(new_updates, new_mu, new_nu)
            return new_updates

        @staticmethod
        def backward(ctx, dupdates):
            new_updates, new_mu, new_nu = ctx.saved_tensors
            b1, b2, eps_root, count = ctx.others
            out = ref_ops.backward_updates(dupdates, new_updates, new_mu, new_nu, b1, b2, eps_root, count)
            return out[0], out[1], None

    class RefAdamOp:  # adam_op.py:124-181
        def __init__(self, b1=0.9, b2=0.999, eps=1e-8, *, eps_root=0.0, inplace=True):
            self.b1, self.b2, self.eps, self.eps_root, self.inplace = b1, b2, eps, eps_root, inplace

        def __call__(self, mu, nu, updates, count):
            if updates is None:
                return mu, nu, None
            if self.inplace:
                new_updates, new_mu, new_nu = ref_ops.forward_(updates, mu, nu, self.b1, self.b2, self.eps,
                                                               self.eps_root, count)
                return new_mu, new_nu, new_updates
            new_mu = MuOp.apply(updates, mu, self.b1)
            new_nu = NuOp.apply(updates, nu, self.b2)
            new_updates = UpdatesOp.apply(new_mu, new_nu, (self.b1, self.b2, self.eps, self.eps_root, count))
            return new_mu, new_nu, new_updates

    return RefAdamOp


class RefMetaAdam:
    """``MetaAdam(net, lr, use_accelerated_op=True)`` as the reference runs it on CPU tensors: per step
    ``autograd.grad(loss, params, create_graph=True)``, then PER LEAF: ``inc_count`` (3 small torch ops on a 0-dim
    int64 tensor, transform/utils.py:111-122), the three-Function AdamOp, ``u.mul(-lr)``, ``p.add(update)``."""

    def __init__(self, module: torch.nn.Module, lr: float = 1e-3, betas: Sequence[float] = (0.9, 0.999),
                 eps: float = 1e-8, *, eps_root: float = 0.0, moment_requires_grad: bool = True) -> None:
        ref = ao.load_ref()
        if ref is None:
            raise RuntimeError("oracle/_ref/_C.so is missing: python oracle/build_ref.py")
        self.op = make_ref_adam_op(ref.adam_op)(betas[0], betas[1], eps, eps_root=eps_root, inplace=False)
        self.lr = lr
        self.slots = [(m, k) for m in module.modules() for k, p in m._parameters.items() if p is not None]
        params = self.params()
        self.mu = [torch.zeros_like(p, requires_grad=moment_requires_grad) for p in params]
        self.nu = [torch.zeros_like(p, requires_grad=moment_requires_grad) for p in params]
        self.count = [torch.zeros((), dtype=torch.int64) for _ in params]
        self.int64_max = torch.iinfo(torch.int64).max

    def params(self) -> list:
        return [m._parameters[k] for m, k in self.slots]

    def step(self, loss: torch.Tensor) -> None:
        params = self.params()
        grads = torch.autograd.grad(loss, params, create_graph=True, allow_unused=True)
        for i, (p, g) in enumerate(zip(params, grads)):
            if g is None:
                continue
            c = self.count[i]
            self.count[i] = c + (c != self.int64_max).to(dtype=c.dtype)          # inc_count
            self.mu[i], self.nu[i], u = self.op(self.mu[i], self.nu[i], g, int(self.count[i]))
            m, k = self.slots[i]
            m._parameters[k] = p.add(u.mul(-self.lr))                              # scale(-lr) ; apply_updates
