#!/usr/bin/env python
"""A differentiable optimizer step over a PERSISTENT state, recorded in ONE CUDA graph (round 2, `capturable=True`).

    python examples/online_meta_gradient.py [--log2n 24] [--steps 50]

The pattern is online hyper-parameter adaptation by one-step look-ahead (meta-gradient RL, learned loss weights, ...):
every iteration takes ONE differentiable Adam step on the model parameters, evaluates an outer objective at the new
parameters, differentiates it w.r.t. a hyper-parameter of the inner loss THROUGH that step, and then continues from the
new parameters with the graph cut.  The optimizer state (moments, step count) persists across iterations -- which is
exactly what could not be recorded in a CUDA graph before: the reference reads `count` back to the host on every call
(torchopt/accelerated_op/adam_op.py:98-101), and this package's host-scalar bias corrections would be frozen at their
recording-time value (it refuses that).  With `FuncAdam(..., capturable=True)` the step counts live on the device, the
bias corrections come from host-computed device tables, and a whole iteration -- inner gradient, differentiable step,
outer loss, meta-backward, parameter write-back, hyper-parameter update -- replays as one graph launch, bit-identical to
eager (tests/test_capturable_gpu.py).

Synthetic problem: parameters p (one flat leaf of n elements), per-element loss weights w (the hyper-parameter, n
elements), inner gradient g = w * p, outer objective sum(p'^2).  Prints one JSON line with the time per iteration for
(a) the default optimizer run eagerly, (b) the capturable optimizer run eagerly, (c) the capturable iteration replayed
from a CUDA graph, and the algorithmic HBM rate of (c).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchopt_b200 as tb  # noqa: E402


def make_iteration(params: torch.Tensor, hyper: torch.Tensor, opt, hyper_lr: float = 1e-2):
    """One online iteration on static buffers `params` / `hyper` (both updated in place)."""

    def iteration():
        leaf = params.detach().requires_grad_(True)
        w = hyper.detach().requires_grad_(True)
        (new_p,) = opt.apply_gradients([w * leaf], [leaf])          # ONE differentiable whole-step launch
        (meta,) = torch.autograd.grad((new_p * new_p).sum(), [w])   # ... and one backward launch through it
        with torch.no_grad():
            params.copy_(new_p)                                      # persistent parameters: static address
            hyper.sub_(hyper_lr * meta)                              # hyper-parameter step
        tb.stop_gradient(opt)                                        # the state continues by value (no-op when capturable)
        return meta

    return iteration


def measure(log2n: int = 24, steps: int = 50) -> dict:
    dev = torch.device("cuda")
    n = 1 << log2n

    def fresh():
        gen = torch.Generator(device=dev).manual_seed(0)
        return (torch.randn(n, device=dev, generator=gen) * 0.3, torch.rand(n, device=dev, generator=gen) + 0.5)

    def timed(fn, iters):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    out = {"workload": f"online one-step look-ahead on one flat leaf of 2^{log2n} fp32 elements", "n": n}
    p, w = fresh()
    out["eager_default_ms"] = timed(make_iteration(p, w, tb.FuncAdam(1e-3, eps_root=1e-8)), steps)
    p, w = fresh()
    out["eager_capturable_ms"] = timed(make_iteration(p, w, tb.FuncAdam(1e-3, eps_root=1e-8, capturable=True)), steps)
    p, w = fresh()
    opt = tb.FuncAdam(1e-3, eps_root=1e-8, capturable=True)
    step = tb.graphs.capture(make_iteration(p, w, opt))
    out["cuda_graph_capturable_ms"] = timed(step, steps)
    # what one iteration moves at the very least: w*p (12 B), step forward (28), moment write-back (16), outer product and
    # its backward (8 + 12), step backward without incoming moment gradients (28), w-gradient (12), copy-back + hyper (28)
    out["cuda_graph_algorithmic_gbs"] = 144 * n / (out["cuda_graph_capturable_ms"] * 1e-3) / 1e9
    st = next(s for s in opt.state if isinstance(s, tb.ScaleByAdamState))
    out["device_step_count_after"] = int(st.count[0])
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=24)
    ap.add_argument("--steps", type=int, default=50)
    a = ap.parse_args()
    print(json.dumps(measure(a.log2n, a.steps)))
