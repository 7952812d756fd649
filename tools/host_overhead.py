#!/usr/bin/env python
"""Where does the time of a 62-leaf (ResNet-18-sized) differentiable step go?  Host-side micro-benchmark."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchopt_b200 as tb
from bench import RESNET18_LEAVES

dev = torch.device("cuda")
L = RESNET18_LEAVES
p = [torch.randn(k, device=dev).requires_grad_(True) for k in L]
g = [torch.randn(k, device=dev) * 1e-2 for k in L]
mu = [torch.zeros(k, device=dev) for k in L]
nu = [torch.zeros(k, device=dev) for k in L]
cnt = [1] * len(L)


def timeit(name, fn, iters=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name:55s} host {1e6*(t1-t0)/iters:8.1f} us/call   incl. drain {1e6*(t2-t0)/iters:8.1f} us/call")


pd = [x.detach() for x in p]
timeit("adam_op.fused_step (62 leaves, no autograd)", lambda: tb.adam_op.fused_step(pd, g, mu, nu, .9, .999, 1e-8, 0., -1e-3, cnt, False, False))
timeit("adam_op.fused_forward (62 leaves)", lambda: tb.adam_op.fused_forward(g, mu, nu, .9, .999, 1e-8, 0., cnt, False))
step = tb.AdamStep(step_size=-1e-3)
timeit("AdamStep()(...) with autograd node (62 leaves)", lambda: step(p, g, mu, nu, cnt))
timeit("62 x torch.empty_like", lambda: [torch.empty_like(x) for x in pd])
timeit("62 x (w * p) torch mul", lambda: [a * b for a, b in zip(g, pd)])
from torchopt_b200.transform import inc_count_flat, make_counts
counts = make_counts([0] * len(L), [dev] * len(L))
timeit("inc_count_flat (62 leaves)", lambda: inc_count_flat(g, counts))
opt = tb.FuncAdam(1e-3)
def full():
    opt.state = None
    cur = list(p)
    for _ in range(5):
        cur = opt.apply_gradients([w * q for w, q in zip(g, cur)], cur)
    outer = torch.stack([q.sum() for q in cur]).sum()
    return torch.autograd.grad(outer, p)
timeit("full 5-step unroll + meta-backward", full, iters=20)
def fwd_only():
    opt.state = None
    cur = list(p)
    for _ in range(5):
        cur = opt.apply_gradients([w * q for w, q in zip(g, cur)], cur)
    return cur
timeit("5-step unroll forward only", fwd_only, iters=20)
timeit("opt.impl.init (62 leaves zeros + counts)", lambda: opt.impl.init(pd), iters=50)


# ---- round 2: the default state layout (flat moment buffers, host counts) vs the per-leaf reference layout ----------
from torchopt_b200 import optim  # noqa: E402
gd = [x.detach() for x in g]


def five_steps(flat):
    optim.FLAT_STATE = flat
    o = tb.FuncAdam(1e-3)
    cur = list(p)
    for _ in range(5):
        cur = o.apply_gradients(gd, cur)
    return cur


def five_steps_and_backward(flat):
    cur = five_steps(flat)
    return torch.autograd.grad(torch.stack([q.sum() for q in cur]).sum(), p)


for flat in (True, False):
    tag = "flat moment buffers (default)" if flat else "per-leaf state (optim.FLAT_STATE = False)"
    timeit(f"5 x FuncAdam.apply_gradients, {tag}", lambda: five_steps(flat), iters=40)
    timeit(f"  + meta-backward (62 sums + stack), {tag}", lambda: five_steps_and_backward(flat), iters=40)
optim.FLAT_STATE = True
