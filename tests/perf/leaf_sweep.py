#!/usr/bin/env python
"""BASELINE config 5: leaf-size sweep 4K..1G elements (fp32 and bf16-grad/fp32-state) and many-small-leaf
multi-tensor launches, this repo's kernels vs the reference's own CUDA adam_op recompiled for sm_100
(oracle/_ref/cuda/_C.so, built from the unmodified sources by `python oracle/build_ref.py --cuda`).

    python tests/perf/leaf_sweep.py --out gpurun_out/leaf_sweep.json    # on the GPU box

"ours"      : fused forward + fused backward through the C ABI (2 launches per step, any number of leaves)
"reference" : what autograd runs per leaf per step with the reference extension: forward_mu, forward_nu,
              forward_updates, backward_updates, 2 accumulation adds, backward_mu, backward_nu, 1 add
              (accelerated_op/adam_op.py:168-180,105-121,53-60,79-86) -- 9 launches per leaf per step.
Timing: CUDA events, 5 warm-ups + 20 iterations; sizes whose working set fits the 126 MB L2 are marked.
This is TEST infrastructure (it lives under tests/ because it loads oracle/_ref/cuda); nothing in the product imports it.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

B1, B2, EPS, ER, CNT = 0.9, 0.999, 1e-8, 1e-8, 3
tb_adam_op = None
RESNET18 = (
    [9408, 64, 64] + [36864, 64, 64] * 4
    + [73728, 128, 128, 147456, 128, 128, 8192, 128, 128] + [147456, 128, 128] * 2
    + [294912, 256, 256, 589824, 256, 256, 32768, 256, 256] + [589824, 256, 256] * 2
    + [1179648, 512, 512, 2359296, 512, 512, 131072, 512, 512] + [2359296, 512, 512] * 2
    + [512000, 1000]
)


def load_ref_cuda():
    path = os.path.join(ROOT, "oracle", "_ref", "cuda", "_C.so")
    if not os.path.isfile(path):
        return None
    import torch  # noqa: F401
    spec = importlib.util.spec_from_file_location("_C", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.adam_op


def time_ms(torch, fn, iters=20, warmup=5, repeats=1):
    """best of ``repeats`` timings of ``iters`` calls (CUDA events on the current stream)"""
    for _ in range(warmup):
        fn()
    best = float("inf")
    for _ in range(repeats):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best


def spin_up(torch, seconds=0.5):
    """Bring the GPU out of its idle power state before the first (smallest, launch-bound) measurement: round 1's
    sweep started cold and its first row (2^12: 30 us) was slower than its third (2^16: 11 us) for that reason only."""
    import time
    a = torch.empty(1 << 26, device="cuda")
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(20):
            a.add_(1.0)
        torch.cuda.synchronize()
    del a


def graph_ms(torch, step, steps_per_graph=20, replays=10):
    """ms per step with ``steps_per_graph`` steps recorded in ONE CUDA graph: no host launch cost at all, i.e. what the
    device itself needs for the back-to-back launches (programmatic edges between them when PDL is on)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step(side.cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(steps_per_graph):
            step(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        g.replay()
    best = float("inf")
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(replays):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / (replays * steps_per_graph))
    return best


def make_leaves(torch, sizes, grad_dtype, state_dtype=None):
    dev = torch.device("cuda")
    sd = state_dtype or torch.float32
    t = {k: [] for k in ("g", "mu", "nu", "du", "dme", "dne", "mu2", "nu2", "u", "dg", "dmu", "dnu")}
    for n in sizes:
        t["g"].append((torch.randn(n, device=dev) * 1e-2).to(grad_dtype))
        t["mu"].append((torch.randn(n, device=dev) * 1e-2).to(sd))
        t["nu"].append((torch.rand(n, device=dev) * 1e-4).to(sd))
        t["du"].append(torch.randn(n, device=dev).to(grad_dtype))
        t["dme"].append((torch.randn(n, device=dev) * 1e-3).to(sd))
        t["dne"].append((torch.randn(n, device=dev) * 1e-3).to(sd))
        for k in ("mu2", "nu2", "dmu", "dnu"):
            t[k].append(torch.empty(n, device=dev, dtype=sd))
        for k in ("u", "dg"):
            t[k].append(torch.empty(n, device=dev, dtype=grad_dtype))
    return t


def ours(torch, abi, t, sizes, dtype, per_leaf=False, graph=False, iters=20, repeats=1):
    ptr = {k: [x.data_ptr() for x in v] for k, v in t.items()}
    stream = torch.cuda.current_stream().cuda_stream

    def prep(idx):
        sel = lambda k: [ptr[k][i] for i in idx]  # noqa: E731
        return abi.PreparedFused(dtype, [sel(k) for k in ("g", "mu", "nu", "mu2", "nu2", "u")],
                                 [sel(k) for k in ("du", "u", "mu2", "g", "dme", "dne", "dg", "dmu", "dnu")],
                                 [sizes[i] for i in idx], [CNT] * len(idx), B1, B2, EPS, ER)

    preps = [prep([i]) for i in range(len(sizes))] if per_leaf else [prep(list(range(len(sizes))))]

    def step(s=stream):
        for p in preps:
            p.forward(s)
        for p in preps:
            p.backward(s)

    if graph:
        return graph_ms(torch, step)
    return time_ms(torch, step, iters=iters, repeats=repeats)


def reference(torch, ref, t, sizes):
    def step():
        for i in range(len(sizes)):
            g, mu, nu = t["g"][i], t["mu"][i], t["nu"][i]
            new_mu = ref.forward_mu(g, mu, B1)
            new_nu = ref.forward_nu(g, nu, B2)
            u = ref.forward_updates(new_mu, new_nu, B1, B2, EPS, ER, CNT)
            dmu_new, dnu_new = ref.backward_updates(t["du"][i], u, new_mu, new_nu, B1, B2, ER, CNT)
            dmu_tot = t["dme"][i] + dmu_new
            dnu_tot = t["dne"][i] + dnu_new
            dg_mu, _ = ref.backward_mu(dmu_tot, g, mu, B1)
            dg_nu, _ = ref.backward_nu(dnu_tot, g, nu, B2)
            _ = dg_mu + dg_nu

    return time_ms(torch, step)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "leaf_sweep.json"))
    ap.add_argument("--max-log2", type=int, default=30)
    args = ap.parse_args()
    import torch

    import torchopt_b200
    from torchopt_b200 import abi
    abi.load()
    global tb_adam_op
    tb_adam_op = torchopt_b200.adam_op
    ref = load_ref_cuda()
    torch.manual_seed(0)
    rows = []
    spin_up(torch)
    for lg in range(12, args.max_log2 + 1, 2):
        n = 1 << lg
        for name, gd, sd, dt, bpe in (("fp32", torch.float32, torch.float32, abi.F32, 60),
                                      ("bf16-grad/fp32-state", torch.bfloat16, torch.float32, abi.BF16_F32, 50),
                                      ("fp64", torch.float64, torch.float64, abi.F64, 120)):
            if name == "fp64" and lg > 28:
                continue    # 12 fp64 arrays of 2^30 elements do not fit next to the allocator's cache
            t = make_leaves(torch, [n], gd, sd)
            small = lg <= 22
            ms = ours(torch, abi, t, [n], dt, iters=200 if small else 20, repeats=3)
            row = {"case": "single leaf", "n": n, "log2n": lg, "dtype": name, "ours_ms": ms,
                   "ours_gelem_s": n / (ms * 1e-3) / 1e9, "ours_gbs": bpe * n / (ms * 1e-3) / 1e9,
                   "fits_l2": 12 * t["g"][0].element_size() * n < 100e6}
            if lg <= 26:
                # the same two launches per step recorded in a CUDA graph (no host launch cost), with and without PDL
                gms = ours(torch, abi, t, [n], dt, graph=True)
                abi.set_pdl(False)
                gms_plain = ours(torch, abi, t, [n], dt, graph=True)
                ms_plain = ours(torch, abi, t, [n], dt, iters=200 if small else 20, repeats=3)
                abi.set_pdl(True)
                row.update(ours_graph_ms=gms, ours_graph_gbs=bpe * n / (gms * 1e-3) / 1e9, ours_graph_no_pdl_ms=gms_plain,
                           ours_no_pdl_ms=ms_plain)
            if gd == torch.float32:
                # the same nine calls per leaf through THIS repo's drop-in `_C.adam_op` surface (no Python changes
                # on the TorchOpt side: INTEGRATION.md section 2 only)
                dms = reference(torch, tb_adam_op, t, [n])
                row.update(dropin_surface_ms=dms, dropin_surface_gelem_s=n / (dms * 1e-3) / 1e9)
            if ref is not None and gd == torch.float32:
                rms = reference(torch, ref, t, [n])
                row.update(ref_cuda_ms=rms, ref_cuda_gelem_s=n / (rms * 1e-3) / 1e9, speedup=rms / ms,
                           dropin_speedup=rms / dms)
            rows.append(row)
            print(json.dumps(row), flush=True)
            del t
            torch.cuda.empty_cache()
    for label, sizes in (("1024 leaves x 4096", [4096] * 1024), ("ResNet-18 leaf set (62 leaves)", list(RESNET18)),
                         ("Omniglot CNN leaf set (14 leaves)", [576, 64, 64, 64] + [36864, 64, 64, 64] * 2 + [320, 5])):
        t = make_leaves(torch, sizes, torch.float32)
        n = sum(sizes)
        ms_mt = ours(torch, abi, t, sizes, abi.F32)
        ms_pl = ours(torch, abi, t, sizes, abi.F32, per_leaf=True)
        row = {"case": label, "n": n, "leaves": len(sizes), "dtype": "fp32", "ours_multi_tensor_ms": ms_mt,
               "ours_multi_tensor_gelem_s": n / (ms_mt * 1e-3) / 1e9, "ours_per_leaf_ms": ms_pl,
               "ours_launches_multi_tensor": 2, "ours_launches_per_leaf": 2 * len(sizes)}
        dms = reference(torch, tb_adam_op, t, sizes)
        row.update(dropin_surface_ms=dms)
        if ref is not None:
            rms = reference(torch, ref, t, sizes)
            row.update(ref_cuda_ms=rms, ref_cuda_launches=9 * len(sizes), speedup=rms / ms_mt, dropin_speedup=rms / dms)
        rows.append(row)
        print(json.dumps(row), flush=True)
        del t
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"reference_cuda_available": ref is not None, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
