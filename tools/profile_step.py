#!/usr/bin/env python
"""Minimal driver for ncu: 3 warm-up + 2 profiled launches of the whole-step forward and backward kernels (fp32,
n = 2^28, one flat leaf).   ncu --set full -k regex:stream_kernel -s 6 -c 2 python tools/profile_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torchopt_b200 import abi

n = 1 << 28
dev = torch.device("cuda")
torch.manual_seed(0)
t = {k: torch.randn(n, device=dev) * sc for k, sc in (("p", .1), ("g", 1e-2), ("mu", 1e-2), ("dp", 1.), ("dme", 1e-3), ("dne", 1e-3))}
t["nu"] = torch.rand(n, device=dev) * 1e-4
o = {k: torch.empty(n, device=dev) for k in ("p2", "mu2", "nu2", "dg", "dmu", "dnu")}
P = {k: [v.data_ptr()] for k, v in {**t, **o}.items()}
s = torch.cuda.current_stream().cuda_stream
for _ in range(4):   # launches 0..7: fwd, bwd alternating; ncu -s 6 -c 2 captures the 4th pair
    abi.step_forward(abi.F32, P["p"], P["g"], P["mu"], P["nu"], P["p2"], P["mu2"], P["nu2"], None, [n], [3], .9, .999, 1e-8, 1e-8, -1e-3, s)
    abi.step_backward(abi.F32, P["dp"], None, P["mu2"], P["nu2"], P["g"], P["dme"], P["dne"], P["dg"], P["dmu"], P["dnu"], [n], [3], .9, .999, 1e-8, 1e-8, -1e-3, s)
torch.cuda.synchronize()
print("profile_step done")
