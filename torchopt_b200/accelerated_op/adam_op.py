"""``AdamOp`` -- host-side dispatch of the accelerated differentiable Adam operator for B200.

Mirrors the public surface of the reference class (/root/reference/torchopt/accelerated_op/adam_op.py:33-181):
``AdamOp(b1, b2, eps, *, eps_root, inplace)(mu, nu, updates, count) -> (new_mu, new_nu, new_updates)`` with
nested ``MuOp`` / ``NuOp`` / ``UpdatesOp`` autograd Functions.  What is different underneath:

* the differentiable path is ONE autograd node (``FusedOp``) whose forward is one fused launch (24 B/elem
  instead of 36) and whose backward is one fused launch (36 B/elem instead of 60-84 incl. the autograd engine's
  accumulation kernels); it saves 3 tensors (g, u, mu') instead of 7;
* ``AdamOp.multi`` runs ALL leaves of a pytree through that one node -- one launch forward, one backward --
  instead of the per-leaf Python loop of the reference (transform/scale_by_adam.py:325-332);
* ``count`` may carry a host mirror (see torchopt_b200.transform) so no device->host sync happens per leaf.

Results are bit-identical to the reference graph: each engine accumulation the fusion absorbs has exactly two
addends (IEEE addition is commutative), see tests/test_unroll_gpu.py.  There is no CPU path: non-CUDA tensors
raise ``RuntimeError('Not implemented')`` from the native module.
"""
from __future__ import annotations

import os
from typing import Any, Sequence

import torch

from torchopt_b200._native import adam_op

INT64_MAX = torch.iinfo(torch.int64).max

# A/B switch for parity tests: True runs the fused differentiable path through the NATIVE autograd node
# (torch_shim.cpp:FusedNode -- the same kernels and saved tensors without ~200 us of interpreter bookkeeping per
# 62-leaf call); False through the Python autograd.Function below (AdamOp.FusedOp), kept as the readable specification.
NATIVE_NODE = os.environ.get("B200_ADAM_PY_NODES", "") in ("", "0")


def capturing() -> bool:
    """True while the current CUDA stream is recording a CUDA graph."""
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


def tag_count(count: torch.Tensor, host: int, in_capture: bool | None = None) -> torch.Tensor:
    """Attach the host mirror of a step counter to its device tensor, together with the tensor's version so that an
    in-place edit of the counter (which bumps the version) invalidates the mirror, and with whether the counter was
    created while a CUDA graph was being recorded (see ``check_capturable``)."""
    count._b200_host_count = host  # type: ignore[attr-defined]
    count._b200_host_version = count._version  # type: ignore[attr-defined]
    count._b200_in_capture = capturing() if in_capture is None else in_capture  # type: ignore[attr-defined]
    return count


def check_capturable(count: Any) -> None:
    """Bias corrections are host scalars baked into the launch, so a recorded step replays with the step count it was
    recorded at.  That is right when the optimizer state is created inside the recorded function (a task's inner loop
    always counts 1..K from a fresh state) and silently WRONG when a persistent state is stepped inside a capture
    (e.g. a recorded training iteration around ``Adam.step()``): refuse the latter."""
    if isinstance(count, torch.Tensor) and not getattr(count, "_b200_in_capture", True) and capturing():
        raise RuntimeError(
            "torchopt_b200: an optimizer state created OUTSIDE a CUDA-graph capture is being stepped INSIDE one. The "
            "bias corrections 1/(1-b^count) are host scalars, so every replay would reuse the step count of the "
            "recording. Create the state inside the captured function (fresh state per replay, as in a MAML inner "
            "loop), or -- for an in-place training step -- build the optimizer with capturable=True (Adam / AdamW / "
            "FlatAdam / FlatAdamW), which keeps the step count on the device.")


def count_as_int(count: Any) -> int:
    """Step count as a Python int.  A tensor tagged by ``torchopt_b200.transform`` answers from its host mirror as long
    as it has not been modified in place since; otherwise (untagged or edited tensor) it is read with ``int()`` -- a
    device->host sync, which is what the reference's pybind ``size_t`` conversion does on EVERY call
    (accelerated_op/adam_op.py:98-101) -- and re-tagged."""
    if isinstance(count, int):
        return count
    host = getattr(count, "_b200_host_count", None)
    if host is not None and getattr(count, "_b200_host_version", None) == count._version:
        return host
    host = int(count)
    if isinstance(count, torch.Tensor):
        tag_count(count, host)
    return host


class AdamOp:  # pylint: disable=too-few-public-methods
    """Fused accelerated Adam operators (B200 build)."""

    # ---- reference-compatible three-node decomposition (kept so code that reaches for AdamOp.MuOp etc.
    # ---- keeps working; not used by __call__ unless fused=False) --------------------------------------
    class MuOp(torch.autograd.Function):  # pylint: disable=abstract-method
        """First-moment update ``mu' = b1*mu + (1-b1)*g`` (reference: adam_op.py:36-60)."""

        @staticmethod
        def forward(ctx: Any, updates: torch.Tensor, mu: torch.Tensor, b1: float) -> torch.Tensor:
            ctx.b1 = b1
            ctx.save_for_backward(updates, mu)
            return adam_op.forward_mu(updates, mu, b1)

        @staticmethod
        def backward(ctx: Any, dmu: torch.Tensor) -> Any:
            updates, mu = ctx.saved_tensors
            dupdates, dmu_prev = adam_op.backward_mu(dmu, updates, mu, ctx.b1)
            return dupdates, dmu_prev, None

    class NuOp(torch.autograd.Function):  # pylint: disable=abstract-method
        """Second-moment update ``nu' = b2*nu + (1-b2)*g*g`` (reference: adam_op.py:62-86)."""

        @staticmethod
        def forward(ctx: Any, updates: torch.Tensor, nu: torch.Tensor, b2: float) -> torch.Tensor:
            ctx.b2 = b2
            ctx.save_for_backward(updates, nu)
            return adam_op.forward_nu(updates, nu, b2)

        @staticmethod
        def backward(ctx: Any, dnu: torch.Tensor) -> Any:
            updates, nu = ctx.saved_tensors
            dupdates, dnu_prev = adam_op.backward_nu(dnu, updates, nu, ctx.b2)
            return dupdates, dnu_prev, None

    class UpdatesOp(torch.autograd.Function):  # pylint: disable=abstract-method
        """Bias-corrected update ``u = mu_hat / (sqrt(nu_hat + eps_root) + eps)`` (reference: adam_op.py:88-121)."""

        @staticmethod
        def forward(ctx: Any, new_mu: torch.Tensor, new_nu: torch.Tensor, others: tuple) -> torch.Tensor:
            b1, b2, eps, eps_root, count = others
            count = count_as_int(count)
            new_updates = adam_op.forward_updates(new_mu, new_nu, b1, b2, eps, eps_root, count)
            ctx.others = (b1, b2, eps_root, count)
            ctx.save_for_backward(new_updates, new_mu, new_nu)
            return new_updates

        @staticmethod
        def backward(ctx: Any, dupdates: torch.Tensor) -> Any:
            new_updates, new_mu, new_nu = ctx.saved_tensors
            b1, b2, eps_root, count = ctx.others
            dmu, dnu = adam_op.backward_updates(dupdates, new_updates, new_mu, new_nu, b1, b2, eps_root, count)
            return dmu, dnu, None

    # ---- the B200 path: one node for any number of leaves --------------------------------------------
    class FusedOp(torch.autograd.Function):  # pylint: disable=abstract-method
        """``(g_i, mu_i, nu_i) for all leaves -> (mu'_i, nu'_i, u_i) for all leaves`` in one launch each way.

        Tensor arguments are laid out ``g_0..g_{L-1}, mu_0..mu_{L-1}, nu_0..nu_{L-1}``; outputs
        ``mu'_0.., nu'_0.., u_0..``.  Backward equals the reference graph (UpdatesOp.backward, engine adds,
        MuOp.backward + NuOp.backward, engine add) -- see include/b200_adam.h:b200_adam_fused_backward.
        """

        @staticmethod
        def forward(ctx: Any, hyper: tuple, counts: tuple, *tensors: torch.Tensor) -> Any:
            b1, b2, eps, eps_root = hyper
            num = len(counts)
            grads, mus, nus = tensors[:num], tensors[num:2 * num], tensors[2 * num:]
            new_mu, new_nu, new_updates = adam_op.fused_forward(
                list(grads), list(mus), list(nus), b1, b2, eps, eps_root, list(counts), False)
            ctx.hyper, ctx.counts = (b1, b2, eps_root), counts
            ctx.set_materialize_grads(False)  # absent gradients stay None -> the kernel skips those streams
            ctx.save_for_backward(*grads, *new_updates, *new_mu)
            return (*new_mu, *new_nu, *new_updates)

        @staticmethod
        def backward(ctx: Any, *douts: torch.Tensor | None) -> Any:
            b1, b2, eps_root = ctx.hyper
            num = len(ctx.counts)
            saved = ctx.saved_tensors
            grads, new_updates, new_mu = saved[:num], saved[num:2 * num], saved[2 * num:]
            dmu_ext, dnu_ext, dupdates = douts[:num], douts[num:2 * num], douts[2 * num:]
            need = ctx.needs_input_grad[2:]
            dg, dmu, dnu = adam_op.fused_backward(
                list(dupdates), list(dmu_ext), list(dnu_ext), list(new_updates), list(new_mu), list(grads),
                list(need[:num]), list(need[num:2 * num]), list(need[2 * num:]), b1, b2, eps_root, list(ctx.counts))
            return (None, None, *dg, *dmu, *dnu)

    # pylint: disable-next=too-many-arguments
    def __init__(self, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8, *, eps_root: float = 0.0,
                 inplace: bool = True, fused: bool = True) -> None:
        """Same hyper-parameters as the reference (adam_op.py:124-138); ``fused=False`` selects the
        reference's three-node graph (three launches per leaf) and exists for A/B parity tests."""
        self.b1, self.b2, self.eps, self.eps_root = b1, b2, eps, eps_root
        self.inplace = inplace
        self.fused = fused

    def __call__(self, mu: torch.Tensor, nu: torch.Tensor, updates: torch.Tensor | None,
                 count: Any) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor | None]:
        """One leaf (reference call signature, adam_op.py:140-181)."""
        if updates is None:
            return mu, nu, None
        count = count_as_int(count)
        if self.inplace:
            new_updates, new_mu, new_nu = adam_op.forward_(updates, mu, nu, self.b1, self.b2, self.eps,
                                                           self.eps_root, count)
            return new_mu, new_nu, new_updates
        if not self.fused:
            new_mu = self.MuOp.apply(updates, mu, self.b1)
            new_nu = self.NuOp.apply(updates, nu, self.b2)
            new_updates = self.UpdatesOp.apply(new_mu, new_nu, (self.b1, self.b2, self.eps, self.eps_root, count))
            return new_mu, new_nu, new_updates
        new_mu, new_nu, new_updates = self._fused((count,), (updates,), (mu,), (nu,))
        return new_mu, new_nu, new_updates

    def _fused(self, counts: Sequence[int], updates: Sequence, mus: Sequence, nus: Sequence) -> Sequence:
        """``mu'_0.., nu'_0.., u_0..`` through one autograd node (native or Python, see NATIVE_NODE)."""
        if NATIVE_NODE:
            return adam_op.fused_autograd(list(updates), list(mus), list(nus), self.b1, self.b2, self.eps,
                                          self.eps_root, list(counts))
        return self.FusedOp.apply((self.b1, self.b2, self.eps, self.eps_root), tuple(counts), *updates, *mus, *nus)

    def multi(self, mus: Sequence[torch.Tensor | None], nus: Sequence[torch.Tensor | None],
              updates: Sequence[torch.Tensor | None], counts: Sequence[Any],
              ) -> tuple[list, list, list]:
        """All leaves at once: one fused launch (and one autograd node) instead of a Python loop over leaves.

        Leaves whose ``updates`` is None pass their moments through and yield a None update (adam_op.py:148-149);
        leaves whose ``mu`` is None yield (None, None, None) (transform/scale_by_adam.py:321-322).
        """
        num = len(updates)
        live = [i for i in range(num) if updates[i] is not None and mus[i] is not None]
        new_mu, new_nu, new_updates = list(mus), list(nus), [None] * num
        if not live:
            return new_mu, new_nu, new_updates
        host_counts = tuple(count_as_int(counts[i]) for i in live)
        if self.inplace:
            out_mu, out_nu, out_u = adam_op.fused_forward(
                [updates[i] for i in live], [mus[i] for i in live], [nus[i] for i in live],
                self.b1, self.b2, self.eps, self.eps_root, list(host_counts), True)
        elif not self.fused:
            triples = [self(mus[i], nus[i], updates[i], c) for i, c in zip(live, host_counts)]
            out_mu, out_nu, out_u = zip(*triples)
        else:
            flat = self._fused(host_counts, [updates[i] for i in live], [mus[i] for i in live],
                               [nus[i] for i in live])
            k = len(live)
            out_mu, out_nu, out_u = flat[:k], flat[k:2 * k], flat[2 * k:]
        for j, i in enumerate(live):
            new_mu[i], new_nu[i], new_updates[i] = out_mu[j], out_nu[j], out_u[j]
        return new_mu, new_nu, new_updates
