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


def time_ms(torch, fn, iters=20, warmup=5):
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def make_leaves(torch, sizes, grad_dtype):
    dev = torch.device("cuda")
    t = {k: [] for k in ("g", "mu", "nu", "du", "dme", "dne", "mu2", "nu2", "u", "dg", "dmu", "dnu")}
    for n in sizes:
        t["g"].append((torch.randn(n, device=dev) * 1e-2).to(grad_dtype))
        t["mu"].append(torch.randn(n, device=dev) * 1e-2)
        t["nu"].append(torch.rand(n, device=dev) * 1e-4)
        t["du"].append(torch.randn(n, device=dev).to(grad_dtype))
        t["dme"].append(torch.randn(n, device=dev) * 1e-3)
        t["dne"].append(torch.randn(n, device=dev) * 1e-3)
        for k in ("mu2", "nu2", "dmu", "dnu"):
            t[k].append(torch.empty(n, device=dev))
        for k in ("u", "dg"):
            t[k].append(torch.empty(n, device=dev, dtype=grad_dtype))
    return t


def ours(torch, abi, t, sizes, dtype, per_leaf=False):
    ptr = {k: [x.data_ptr() for x in v] for k, v in t.items()}
    stream = torch.cuda.current_stream().cuda_stream

    def prep(idx):
        sel = lambda k: [ptr[k][i] for i in idx]  # noqa: E731
        return abi.PreparedFused(dtype, [sel(k) for k in ("g", "mu", "nu", "mu2", "nu2", "u")],
                                 [sel(k) for k in ("du", "u", "mu2", "g", "dme", "dne", "dg", "dmu", "dnu")],
                                 [sizes[i] for i in idx], [CNT] * len(idx), B1, B2, EPS, ER)

    preps = [prep([i]) for i in range(len(sizes))] if per_leaf else [prep(list(range(len(sizes))))]

    def step():
        for p in preps:
            p.forward(stream)
        for p in preps:
            p.backward(stream)

    return time_ms(torch, step)


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
    for lg in range(12, args.max_log2 + 1, 2):
        n = 1 << lg
        for name, gd, dt, bpe in (("fp32", torch.float32, abi.F32, 60), ("bf16-grad/fp32-state", torch.bfloat16, abi.BF16_F32, 50)):
            t = make_leaves(torch, [n], gd)
            ms = ours(torch, abi, t, [n], dt)
            row = {"case": "single leaf", "n": n, "log2n": lg, "dtype": name, "ours_ms": ms,
                   "ours_gelem_s": n / (ms * 1e-3) / 1e9, "ours_gbs": bpe * n / (ms * 1e-3) / 1e9,
                   "fits_l2": 12 * 4 * n < 100e6}
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
