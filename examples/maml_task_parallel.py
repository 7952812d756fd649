#!/usr/bin/env python
"""BASELINE config 4: MAML few-shot meta-batch, task-parallel over the GPUs of one box.

    python examples/maml_task_parallel.py --iters 3 --check                                   # 1 GPU
    python examples/maml_task_parallel.py --iters 4 --check --graph                           # 1 GPU, CUDA-graph replay per task
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
           --master-port 29533 examples/maml_task_parallel.py --iters 3 --check               # 2 GPUs

Shapes follow /root/reference/examples/few-shot/maml_omniglot.py:64-113 (5-way, 5-shot support = [25,1,28,28],
15-shot query = [75,1,28,28], 32-task meta-batch, the 3-conv-block CNN with 75 205 parameters in 14 leaves, 5 inner
steps, outer torch.optim.Adam(lr=1e-3)); images and labels are synthetic.  The inner optimizer is
``torchopt_b200.MetaAdam`` (the accelerated differentiable Adam: one fused launch per inner step, one per backward step)
in place of the example's MetaSGD so that the hot path is exercised.

Parallelisation: meta-parameters replicated; rank r runs tasks task_slice(32, r, N); the unrolled
inner loops never communicate; ONE NCCL all-reduce of ONE flattened fp32 bucket sums the meta-gradients; every rank applies
the same outer step.  ``--check`` makes every rank also compute the full 32-task meta-gradient alone and compares.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchopt_b200 as tb  # noqa: E402
from torchopt_b200.parallel import GradBucket, task_slice  # noqa: E402


def make_net(n_way: int) -> nn.Module:
    def block(cin):
        return [nn.Conv2d(cin, 64, 3), nn.BatchNorm2d(64, momentum=1.0, affine=True), nn.ReLU(), nn.MaxPool2d(2, 2)]
    return nn.Sequential(*block(1), *block(64), *block(64), nn.Flatten(), nn.Linear(64, n_way))


class TaskRunner:
    """Runs the inner loop of one task on a module whose parameter slots are re-bound to the meta-parameters."""

    def __init__(self, net: nn.Module, inner_steps: int, lr: float, streams: int = 1):
        self.net, self.inner_steps, self.lr, self.streams = net, inner_steps, lr, max(1, streams)
        self.slots = [(m, k) for m in net.modules() for k, p in m._parameters.items() if p is not None]
        self.meta_params = [m._parameters[k] for m, k in self.slots]
        self.bucket = GradBucket(self.meta_params)   # every .grad is a view of ONE flat buffer
        self.task_graphs: list = []
        self.buckets: list = []

    def rebind(self):
        for (m, k), p in zip(self.slots, self.meta_params):
            m._parameters[k] = p

    def query_loss(self, x_spt, y_spt, x_qry, y_qry) -> torch.Tensor:
        self.rebind()
        inner = tb.MetaAdam(self.net, lr=self.lr, moment_requires_grad=True)   # fresh Adam state per task
        for _ in range(self.inner_steps):
            inner.step(F.cross_entropy(self.net(x_spt), y_spt))
        loss = F.cross_entropy(self.net(x_qry), y_qry)
        self.rebind()
        return loss

    def meta_grads(self, data, tasks, num_tasks) -> list:
        self.bucket.zero()
        x_spt, y_spt, x_qry, y_qry = data
        for t in tasks:
            (self.query_loss(x_spt[t], y_spt[t], x_qry[t], y_qry[t]) / num_tasks).backward()
        return [p.grad for p in self.meta_params]

    def meta_grads_graphed(self, data, tasks, num_tasks) -> list:
        """Same as ``meta_grads`` with one task's whole computation (5 inner MetaAdam steps, query loss, meta-backward
        accumulating into the bucket) recorded ONCE in a CUDA graph and replayed per task: 4 small input copies +
        one graph launch per task instead of ~2 500 eager kernel launches and their autograd bookkeeping.  With
        ``streams`` > 1 that many independently captured instances replay side by side (tasks are independent; one
        task alone keeps only a few SMs busy), each into its own flat gradient buffer; the buffers are then summed."""
        x_spt, y_spt, x_qry, y_qry = data
        if not self.task_graphs:
            def one_task(xs, ys, xq, yq):
                (self.query_loss(xs, ys, xq, yq) / num_tasks).backward()
            # instance j accumulates into its own flat gradient buffer (the .grad views attached while it is recorded)
            self.buckets = [self.bucket] + [GradBucket(self.meta_params) for _ in range(self.streams - 1)]
            for b in self.buckets:
                self.task_graphs.append(tb.graphs.capture(one_task, x_spt[0], y_spt[0], x_qry[0], y_qry[0],
                                                          before_each=b.zero))
        for b in self.buckets:
            b.zero()
        self.bucket.attach()
        tb.graphs.replay_concurrently(self.task_graphs, ((x_spt[t], y_spt[t], x_qry[t], y_qry[t]) for t in tasks))
        for b in self.buckets[1:]:           # fixed summation order: deterministic result
            for flat, other in zip(self.bucket.buffers.values(), b.buffers.values()):
                flat.add_(other)
        return [p.grad for p in self.meta_params]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tasks", type=int, default=32)
    ap.add_argument("--inner-steps", type=int, default=5)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--graph", action="store_true", help="replay one captured CUDA graph per task (torchopt_b200.graphs)")
    ap.add_argument("--streams", type=int, default=1, help="with --graph: captured instances replayed concurrently")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    torch.manual_seed(1)                       # same meta-parameters and the same synthetic meta-batch on all ranks
    net = make_net(5).to(dev)
    gen = torch.Generator(device=dev).manual_seed(2)
    T = args.tasks
    data = (torch.randn(T, 25, 1, 28, 28, device=dev, generator=gen), torch.randint(0, 5, (T, 25), device=dev, generator=gen),
            torch.randn(T, 75, 1, 28, 28, device=dev, generator=gen), torch.randint(0, 5, (T, 75), device=dev, generator=gen))
    runner = TaskRunner(net, args.inner_steps, lr=1e-3, streams=args.streams)
    meta_opt = torch.optim.Adam(runner.meta_params, lr=1e-3)
    mine = task_slice(T, rank, world)
    n_params = sum(p.numel() for p in runner.meta_params)

    max_rel = 0.0
    times = []
    for it in range(args.iters):
        if args.check:
            full = [g.clone() for g in runner.meta_grads(data, range(T), T)]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        l0 = tb.abi.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h0 = time.perf_counter()
        if args.graph:
            grads = runner.meta_grads_graphed(data, mine, T)
        else:
            grads = runner.meta_grads(data, mine, T)      # local: sum over my tasks of d(qry_loss / T)
        runner.bucket.allreduce()                          # ONE all-reduce, in place on the flat bucket (no copies)
        host_ms = (time.perf_counter() - h0) * 1e3         # host time to ISSUE the iteration (no synchronisation)
        e1.record()
        torch.cuda.synchronize()
        launches = tb.abi.kernel_launches() - l0
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        times.append(float(ms))
        if args.check:
            for g, f in zip(grads, full):
                max_rel = max(max_rel, float((g - f).abs().max() / f.abs().max().clamp_min(1e-30)))
        meta_opt.step()
    if args.check:
        # identical meta-parameters on every rank after the outer steps
        flat = torch.cat([p.detach().reshape(-1) for p in runner.meta_params])
        if world > 1:
            lo, hi = flat.clone(), flat.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            assert torch.equal(lo, hi), "meta-parameters diverged across ranks"
        assert max_rel < 1e-4, f"task-parallel meta-gradient differs from the single-process one: {max_rel}"
    if rank == 0:
        best = min(times[1:] or times)
        print(json.dumps({"workload": "MAML 5-way 5-shot, 32-task meta-batch, 5 inner MetaAdam steps, 75205-param CNN",
                          "n_gpus": world, "tasks_per_rank": len(mine), "params": n_params, "cuda_graph": args.graph, "streams": args.streams if args.graph else 1,
                          "ms_per_meta_iter": best, "host_issue_ms_last_iter": host_ms, "tasks_per_s": T / (best * 1e-3),
                          "adam_kernel_launches_per_rank_per_iter": launches,
                          "max_rel_diff_vs_single_process": max_rel if args.check else None}))
    if world > 1:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
