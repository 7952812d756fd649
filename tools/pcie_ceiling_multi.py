#!/usr/bin/env python
"""Aggregate host<->device copy throughput when ALL ranks copy at the same time (torchrun, one rank per GPU): the
ceiling of bench.py's `e2e` leg at N > 1.  Each rank moves 6 x 256 MiB in and 6 x 256 MiB out per repetition in
16 MiB slices on two streams (the e2e access pattern without any compute), all ranks between two barriers.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_ceiling_multi.py
"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
m, sl = 1 << 26, 1 << 22
hin = [torch.empty(m, dtype=torch.float32).pin_memory() for _ in range(6)]
hout = [torch.empty(m, dtype=torch.float32).pin_memory() for _ in range(6)]
dbuf = [torch.zeros(sl, dtype=torch.float32, device=dev) for _ in range(12)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def sliced():
    for off in range(0, m, sl):
        with torch.cuda.stream(s1):
            for a in range(6):
                dbuf[a].copy_(hin[a][off:off + sl], non_blocking=True)
        with torch.cuda.stream(s2):
            for a in range(6):
                hout[a][off:off + sl].copy_(dbuf[6 + a], non_blocking=True)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(device_ids=[local])
    torch.cuda.synchronize()


sliced()
reps = 6
barrier()
t0 = time.perf_counter()
for _ in range(reps):
    sliced()
barrier()
dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
each = 6 * 4 * m * reps / float(dt.item()) / 1e9
if rank == 0:
    print(json.dumps({"n_gpus": world, "per_gpu_each_direction_GBps": each, "aggregate_both_directions_GBps": 2 * each * world,
                      "e2e_ceiling_gelem_per_s": each * world / 24.0}))
if world > 1:
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()
