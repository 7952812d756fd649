#!/usr/bin/env python
"""Minimal driver for ncu: the fp64 instantiation of the fused forward / backward (the reference's CUDA tests are
fp64-only, tests/test_alias.py:553-624).  4 forward/backward pairs on one flat leaf of n = 2^26 doubles (arrays of
512 MiB >> L2); `ncu -s 6 -c 2` captures the 4th pair.

    ncu --set full --metrics sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum \
        --clock-control none --import-source on -k regex:stream_kernel -s 6 -c 2 -o gpurun_out/prof_r02_f64 \
        python tools/profile_f64.py
Without ncu it prints CUDA-event timings of the same launches (one JSON line)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torchopt_b200 import abi

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 26)
dev = torch.device("cuda")
torch.manual_seed(0)
dt = torch.float64
t = {k: (torch.randn(n, device=dev) * sc).to(dt) for k, sc in (("g", 1e-2), ("mu", 1e-2), ("du", 1.), ("dme", 1e-3), ("dne", 1e-3))}
t["nu"] = (torch.rand(n, device=dev) * 1e-4).to(dt)
o = {k: torch.empty(n, device=dev, dtype=dt) for k in ("mu2", "nu2", "u", "dg", "dmu", "dnu")}
P = {k: v.data_ptr() for k, v in {**t, **o}.items()}
prep = abi.PreparedFused(abi.F64, [[P[k]] for k in ("g", "mu", "nu", "mu2", "nu2", "u")],
                         [[P[k]] for k in ("du", "u", "mu2", "g", "dme", "dne", "dg", "dmu", "dnu")], [n], [3], .9, .999,
                         1e-8, 1e-8)
s = torch.cuda.current_stream().cuda_stream
ev = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
ev[0].record()
for k in range(4):   # launches 0..7: fwd, bwd alternating
    prep.forward(s)
    ev[2 * k + 1].record()
    prep.backward(s)
    ev[2 * k + 2].record()
torch.cuda.synchronize()
fwd = min(ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(1, 4))
bwd = min(ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) for k in range(1, 4))
print(json.dumps({"n": n, "dtype": "f64", "fwd_ms": fwd, "bwd_ms": bwd, "fwd_GBps": 48 * n / fwd / 1e6,
                  "bwd_GBps": 72 * n / bwd / 1e6, "geometry_fwd": abi.query_launch(abi.F64, False, n),
                  "geometry_bwd": abi.query_launch(abi.F64, True, n)}))
