"""CUDA-graph capture of launch-bound unrolls (the B200 answer to per-leaf / per-step host overhead).

A differentiable K-step inner loop on a small pytree is bound by the host, not by HBM: BASELINE config 2 (ResNet-18
leaf set, 5-step unroll + meta-backward) needs ~0.6 ms of kernel time but ~11 ms of eager PyTorch dispatch, and one
MAML task of config 4 (75 205 parameters, 14 leaves) spends >95 % of its 24 ms in Python / autograd bookkeeping.  The
reference answers nothing here (its per-leaf Python loop, transform/scale_by_adam.py:325-332, is the cost).  On B200
the idiomatic fix is to record the whole computation -- inner forward passes, ``autograd.grad(create_graph=True)``, the
fused Adam step launches, the outer loss and the meta-backward -- ONCE into a CUDA graph and replay it:

    step = capture(meta_grad_fn, x_spt, y_spt, x_qry, y_qry)      # warm-up on a side stream, then capture
    for task in tasks:
        out = step(*task)                                           # 4 small copies + ONE cudaGraphLaunch

Why this works for the Adam path without special cases:

* the leaf table of every fused launch travels in the kernel parameters, so there is no host-built
  device table to go stale between replays;
* bias corrections are host scalars derived from the host-mirrored step counters; a task's inner loop always runs
  counts 1..K from a fresh state, so the recorded scalars are right for every replay;
* nothing on the path synchronises with the host (no ``count.item()``), which capture would reject;
* ``GradBucket`` keeps every ``.grad`` a view of one static buffer, so the captured ``backward()`` accumulates in
  place into memory the caller can read, all-reduce and zero between replays.

Persistent optimizer state: a recorded training iteration around ``Adam.step()`` would replay the bias corrections
of the recording step, so that raises (``accelerated_op.adam_op.check_capturable``) -- unless the optimizer was built
with ``capturable=True`` (``Adam`` / ``AdamW`` / ``FlatAdam`` / ``FlatAdamW``): then the step count lives on the device,
the corrections come from host-computed tables, and every replay is one real, bit-exact optimizer step
(``b200_adam_step_inplace_capturable``, tests/test_capturable_gpu.py).

Python side effects of ``fn`` (building optimizer objects, re-binding module parameters) run during warm-up and
capture only; a replay re-executes the recorded GPU work on the static inputs and nothing else.
"""
from __future__ import annotations

import gc
from typing import Any, Callable

import torch

__all__ = ["CapturedCallable", "capture", "replay_concurrently"]


class CapturedCallable:
    """``fn(*args)`` recorded in one CUDA graph.

    ``args`` are example inputs: CUDA tensors are cloned into static input slots (replays read those), everything
    else is passed through unchanged at capture time and must not change afterwards.  The call returns the *static*
    output tensors of the capture (same objects every time; clone them to keep a value across replays).

    ``fn`` may use autograd freely -- ``autograd.grad(..., create_graph=True)``, ``loss.backward()`` into existing
    ``.grad`` buffers -- but must not synchronise with the host (``.item()``, ``.cpu()``, printing a CUDA tensor).
    """

    def __init__(self, fn: Callable, *args: Any, warmup: int = 3, pool: Any = None,
                 before_each: Callable[[], None] | None = None) -> None:
        if not torch.cuda.is_available():
            raise RuntimeError("torchopt_b200.graphs: CUDA graphs need a CUDA device (there is no CPU path)")
        self.fn = fn
        self.static_inputs = [a.detach().clone() if isinstance(a, torch.Tensor) and a.is_cuda else a for a in args]
        self.before_each = before_each
        # Every instance owns ONE stream for warm-up, capture and (in replay_concurrently) replay: library workspaces
        # (cuBLAS keeps one per handle x stream) recorded in this graph are then never shared with another instance
        # that may be replaying at the same time.
        # Warm-up first: cuDNN / cuBLAS handles, workspaces and autotuning, lazy module loading of every kernel, and
        # the allocator's steady state must all exist before the recording starts.
        # Autograd graphs of earlier eager runs that are only kept alive by unreachable Python cycles would also keep the
        # leaves' gradient-accumulator nodes alive -- bound to the stream they were created on (normally the default
        # stream); the recorded backward would then have to synchronise with that stream, which a capture forbids.
        gc.collect()
        side = self.stream = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        # Leaf parameters created on the default stream keep their gradient-accumulator node there; autograd warns
        # about that on every backward run from another stream.  The stream hand-over is explicit here (wait_stream
        # before, synchronize after, capture on the same private stream), so the warning is noise for this block.
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)
        try:
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    if before_each is not None:
                        before_each()
                    out = fn(*self.static_inputs)
                    del out
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            if before_each is not None:
                before_each()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, pool=pool, stream=self.stream):
                self.outputs = fn(*self.static_inputs)
        finally:
            if quiet is not None:
                quiet(True)
        self.replays = 0

    def pool(self) -> Any:
        """Memory-pool handle of this graph (pass it as ``pool=`` to capture further graphs into the same pool;
        only for graphs that are never replayed concurrently)."""
        return self.graph.pool()

    def load(self, *args: Any) -> None:
        """Copy new inputs into the static slots (tensors that already ARE the slots are skipped)."""
        if len(args) != len(self.static_inputs):
            raise TypeError(f"captured callable takes {len(self.static_inputs)} arguments, got {len(args)}")
        for slot, a in zip(self.static_inputs, args):
            if isinstance(slot, torch.Tensor) and slot.is_cuda:
                if a is not slot:
                    slot.copy_(a, non_blocking=True)
            elif a is not slot and a != slot:
                raise ValueError("non-tensor arguments are baked into the capture and cannot change")

    def replay(self) -> Any:
        """Re-run the recorded work on whatever the static inputs hold now."""
        self.graph.replay()
        self.replays += 1
        return self.outputs

    def __call__(self, *args: Any) -> Any:
        if args:
            self.load(*args)
        return self.replay()


def capture(fn: Callable, *args: Any, warmup: int = 3, pool: Any = None,
            before_each: Callable[[], None] | None = None) -> CapturedCallable:
    """Record ``fn(*args)`` into a CUDA graph after ``warmup`` eager runs; see ``CapturedCallable``.
    ``before_each`` runs (eagerly, outside the graph) before every warm-up run and once before the capture --
    e.g. ``bucket.zero`` so that warm-up runs do not pollute accumulated gradients."""
    return CapturedCallable(fn, *args, warmup=warmup, pool=pool, before_each=before_each)


def replay_concurrently(instances: list, arg_tuples: Any) -> int:
    """Replay independently captured instances of the same computation side by side, each on its own stream.

    A MAML task on a 75 k-parameter network is ~2 500 tiny kernels in one dependency chain: a single replay keeps a
    handful of the 148 SMs busy and is bound by kernel-to-kernel latency.  Tasks are independent, so J instances
    (each captured with its OWN static inputs, memory pool and gradient buffer -- e.g. one ``GradBucket`` each) run
    concurrently: call ``k`` goes to instance ``k mod J``.  The caller's stream waits for all of them on return;
    results live in each instance's static outputs / side-effect buffers (combine them afterwards).
    Returns the number of replays issued."""
    if not instances:
        raise ValueError("replay_concurrently needs at least one captured instance")
    main = torch.cuda.current_stream()
    for inst in instances:
        inst.stream.wait_stream(main)
    issued = 0
    for k, args in enumerate(arg_tuples):
        inst = instances[k % len(instances)]
        with torch.cuda.stream(inst.stream):
            inst(*args)
        issued += 1
    for inst in instances:
        main.wait_stream(inst.stream)
    return issued
