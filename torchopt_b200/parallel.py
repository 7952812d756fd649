"""Multi-GPU plumbing for the Adam hot path (one process per GPU, ``torch.distributed``).

Two cases, and only these two:

* **Flat Adam state** shards naturally: element i depends only on element i of the other arrays, so rank r owns the
  contiguous range ``shard_range(n, r, N)`` (vector-aligned split) or a bin of whole leaves
  (``partition_leaves``) and runs the same kernels on it.  No collective exists on this path.
* **Task-batched meta-learning** (MAML): meta-parameters are replicated, rank r runs the tasks
  ``task_slice(num_tasks, r, N)``, the unrolled inner loops are entirely local, and the meta-gradients are summed
  with ONE all-reduce over ONE flattened fp32 bucket (``allreduce_meta_grads``).  This replaces the reference's RPC
  fan-out + CPU ``torch.mean(torch.stack(...))`` + dist-autograd gradient pull
  (/root/reference/torchopt/distributed/api.py:264-271,315-322, distributed/autograd.py:79-91).
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int, align: int = 8) -> tuple[int, int]:
    """Contiguous element range ``[lo, hi)`` of rank ``rank`` when ``n`` elements are split over ``world`` ranks.

    Interior boundaries are multiples of ``align`` elements (8 fp32 = one 256-bit vector) so every shard base
    stays vector-aligned; the ranges tile ``[0, n)`` exactly and differ in size by less than ``2 * align``.
    """
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world: {rank}/{world}")
    if n < 0 or align <= 0:
        raise ValueError("n must be >= 0 and align > 0")
    blocks = (n + align - 1) // align
    per, extra = divmod(blocks, world)

    def edge(r: int) -> int:
        return min(n, (r * per + min(r, extra)) * align)

    return edge(rank), edge(rank + 1)


def partition_leaves(sizes: Sequence[int], world: int) -> list[list[int]]:
    """Leaf-granular partition: greedy longest-processing-time bin packing of leaf indices over ``world`` ranks.
    Deterministic (ties broken by index), so every rank computes the same assignment without communicating."""
    if world <= 0:
        raise ValueError("world must be positive")
    bins: list[list[int]] = [[] for _ in range(world)]
    load = [0] * world
    for i in sorted(range(len(sizes)), key=lambda k: (-sizes[k], k)):
        r = min(range(world), key=lambda k: (load[k], k))
        bins[r].append(i)
        load[r] += sizes[i]
    for b in bins:
        b.sort()
    return bins


def task_slice(num_tasks: int, rank: int, world: int) -> range:
    """Tasks of a meta-batch owned by ``rank`` (contiguous blocks; sizes differ by at most one)."""
    per, extra = divmod(num_tasks, world)
    lo = rank * per + min(rank, extra)
    return range(lo, lo + per + (1 if rank < extra else 0))


def allreduce_meta_grads(grads: Sequence[torch.Tensor | None], *, scale: float | None = None,
                         group: dist.ProcessGroup | None = None) -> list[torch.Tensor | None]:
    """Sum meta-gradients over ranks with a single all-reduce of one flattened bucket per (device, dtype), in place.

    ``scale`` (e.g. ``1 / num_tasks``) is applied BEFORE the reduction, reproducing the reference's
    ``torch.mean(torch.stack(task_losses))`` (examples/few-shot/maml_omniglot.py:175) when every rank contributes
    the plain sum of its tasks' gradients.  ``None`` entries are treated as zeros of unknown shape and must be
    ``None`` on every rank.  Without an initialised process group this is a local scale only (world size 1).
    """
    live = [g for g in grads if g is not None]
    if not live:
        return list(grads)
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if not multi:                       # world size 1: nothing to exchange, no flatten / copy-back
        if scale is not None:
            for g in live:
                g.mul_(scale)
        return list(grads)
    # one bucket per (device, dtype) -- normally exactly one; fp64 meta-gradients are reduced in fp64, not truncated
    groups: dict = {}
    for g in live:
        groups.setdefault((g.device, g.dtype), []).append(g)
    for members in groups.values():
        bucket = torch.cat([g.reshape(-1) for g in members])
        if scale is not None:
            bucket.mul_(scale)
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for g in members:
            k = g.numel()
            g.copy_(bucket[off:off + k].view_as(g))
            off += k
    return list(grads)


class GradBucket:
    """Flat gradient storage for a parameter list: every ``p.grad`` is a view into ONE buffer per (device, dtype), so
    the meta-gradient all-reduce needs no flatten / copy-back ("makes the meta-grad all-reduce
    bucket free").  ``backward()`` accumulates in place into the views (autograd's AccumulateGrad adds into an existing
    ``.grad``); ``allreduce`` then issues exactly one collective per buffer on the buffer itself.

        bucket = GradBucket(meta_params)
        bucket.zero()                      # instead of optimizer.zero_grad()
        ... (loss / num_tasks).backward() for the local tasks ...
        bucket.allreduce()                 # ONE all_reduce(SUM), in place; .grad views now hold the global gradient
        meta_optimizer.step()
    """

    def __init__(self, params: Sequence[torch.Tensor]) -> None:
        self.params = [p for p in params]
        self.buffers: dict = {}
        self.views: list = [None] * len(self.params)
        groups: dict = {}
        for i, p in enumerate(self.params):
            groups.setdefault((p.device, p.dtype), []).append(i)
        for key, idx in groups.items():
            align = max(1, 128 // self.params[idx[0]].element_size())
            offsets, total = [], 0
            for i in idx:
                total = (total + align - 1) // align * align
                offsets.append(total)
                total += self.params[i].numel()
            flat = torch.zeros(total, device=key[0], dtype=key[1])
            self.buffers[key] = flat
            for i, off in zip(idx, offsets):
                p = self.params[i]
                self.views[i] = flat.as_strided(p.shape, p.stride() if p.is_contiguous() else
                                                torch.empty_like(p).stride(), off)
        self.attach()

    def attach(self) -> None:
        """(Re-)install the views as ``.grad`` (needed after ``zero_grad(set_to_none=True)``)."""
        for p, v in zip(self.params, self.views):
            p.grad = v

    def zero(self) -> None:
        for flat in self.buffers.values():
            flat.zero_()
        self.attach()

    def allreduce(self, *, scale: float | None = None, group: dist.ProcessGroup | None = None) -> None:
        """In-place SUM over ranks of every flat buffer (one collective each; normally exactly one)."""
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        for flat in self.buffers.values():
            if scale is not None:
                flat.mul_(scale)
            if multi:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
