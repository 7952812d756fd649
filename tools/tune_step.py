#!/usr/bin/env python
"""Sweep of the whole-step backward launch shape (unroll x CTA size); build on the CPU box, run on the GPU box.

    python tools/tune_step.py --build ; python tools/tune_step.py --run
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT_DIR = os.path.join(ROOT, "build", "tune_step")
VARIANTS = [(1, 512), (1, 256), (1, 128), (2, 256), (2, 128), (1, 1024)]


def lib_path(u, b):
    return os.path.join(OUT_DIR, f"libb200adam_stepU{u}B{b}.so")


def build_one(v):
    u, b = v
    csrc = os.path.join(ROOT, "torchopt_b200", "csrc")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                           "--fmad=false", "-Xcompiler", "-fPIC", "-shared", "-DB200_ADAM_TUNE_F32_ONLY",
                           f"-DB200_STEP_U_BWD={u}", f"-DB200_STEP_BLOCK_BWD={b}", f"-I{os.path.join(ROOT, 'include')}",
                           f"-I{csrc}", os.path.join(csrc, "b200_adam.cu"), os.path.join(csrc, "host_pipeline.cu"),
                           "-o", lib_path(u, b)])
    return v


def build():
    os.makedirs(OUT_DIR, exist_ok=True)
    with ThreadPoolExecutor(max_workers=6) as ex:
        for v in ex.map(build_one, VARIANTS):
            print("built", v, flush=True)


def run(n_log2, out_path):
    import torch
    from torchopt_b200 import abi
    n = 1 << n_log2
    dev = torch.device("cuda")
    torch.manual_seed(0)
    t = {k: torch.randn(n, device=dev) * sc for k, sc in (("p", .1), ("g", 1e-2), ("mu", 1e-2), ("dp", 1.), ("dme", 1e-3), ("dne", 1e-3))}
    t["nu"] = torch.rand(n, device=dev) * 1e-4
    o = {k: torch.empty(n, device=dev) for k in ("p2", "mu2", "nu2", "dg", "dmu", "dnu", "dpo")}
    P = {k: abi._ptrs([v.data_ptr()]) for k, v in {**t, **o}.items()}
    N, C = abi._arr(abi.c_int64, [n]), abi._arr(abi.c_uint64, [3])
    s = torch.cuda.current_stream().cuda_stream
    rows = []
    for u, b in VARIANTS:
        if not os.path.isfile(lib_path(u, b)):
            continue
        lib = abi.load(lib_path(u, b))
        assert lib.b200_adam_step_forward(abi.F32, 1, P["p"], P["g"], P["mu"], P["nu"], P["p2"], P["mu2"], P["nu2"], None,
                                          N, C, .9, .999, 1e-8, 1e-8, -1e-3, 0., 0., 0, s) == 0
        row = {"U": u, "block": b}
        for name, wd, dpo, nbytes in (("bwd", 0.0, None, 36 * n), ("bwd_l2_decay", 1e-2, P["dpo"], 44 * n)):
            def fn():
                return lib.b200_adam_step_backward(abi.F32, 1, P["dp"], None, P["mu2"], P["nu2"], P["g"], P["dme"],
                                                   P["dne"], P["dg"], P["dmu"], P["dnu"], P["p"], dpo, N, C, .9, .999,
                                                   1e-8, 1e-8, -1e-3, wd, 0., 0, s)
            for _ in range(3):
                assert fn() == 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            row[f"{name}_ms"], row[f"{name}_gbs"] = ms, nbytes / (ms * 1e-3) / 1e9
        rows.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    json.dump({"n": n, "rows": rows}, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--run", action="store_true")
    ap.add_argument("--n-log2", type=int, default=28)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "tune_step.json"))
    a = ap.parse_args()
    if a.build:
        build()
    if a.run:
        run(a.n_log2, a.out)
