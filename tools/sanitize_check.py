#!/usr/bin/env python
"""Small, fast driver for `compute-sanitizer --tool memcheck`: every kernel family over ragged tails, unaligned views,
NULL patterns and multi-leaf tables, with guard elements around every buffer checked afterwards.

    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_check.py
"""
from __future__ import annotations

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from torchopt_b200 import abi  # noqa: E402

GUARD = 64
B1, B2, EPS, ER = 0.9, 0.999, 1e-8, 1e-8


class Arena:
    """One buffer per array with GUARD sentinel elements on both sides of the live range [GUARD+off, GUARD+off+n)."""

    def __init__(self, n, off, dtype):
        self.buf = torch.full((n + 2 * GUARD + off,), 12345.0, device="cuda", dtype=dtype)
        self.lo, self.n = GUARD + off, n
        self.view = self.buf[self.lo:self.lo + n]
        self.view.copy_(torch.rand(n, device="cuda").to(dtype) * 1e-2 + 1e-3)

    def ptr(self):
        return self.view.data_ptr() if self.n else None

    def check(self):
        ok = bool((self.buf[:self.lo] == 12345.0).all()) and bool((self.buf[self.lo + self.n:] == 12345.0).all())
        assert ok, "a kernel wrote outside its range"


def run(dtype_id, tdt, gdt, sizes, off):
    s = torch.cuda.current_stream().cuda_stream
    A = {k: [Arena(n, off, gdt if k in ("p", "p2", "g", "du", "u", "dg") else tdt) for n in sizes]
         for k in ("p", "g", "mu", "nu", "du", "dme", "dne", "mu2", "nu2", "u", "dg", "dmu", "dnu", "p2")}
    P = {k: [a.ptr() for a in v] for k, v in A.items()}
    cnt = [1 + i % 3 for i in range(len(sizes))]
    abi.fused_forward(dtype_id, P["g"], P["mu"], P["nu"], P["mu2"], P["nu2"], P["u"], sizes, cnt, B1, B2, EPS, ER, s)
    abi.fused_backward(dtype_id, P["du"], P["u"], P["mu2"], P["g"], P["dme"], P["dne"], P["dg"], P["dmu"], P["dnu"],
                       sizes, cnt, B1, B2, ER, s)
    none = [None] * len(sizes)
    abi.fused_backward(dtype_id, P["du"], P["u"], P["mu2"], P["g"], none, none, P["dg"], P["dmu"], P["dnu"], sizes, cnt,
                       B1, B2, ER, s)
    abi.fused_backward(dtype_id, none, P["u"], P["mu2"], P["g"], P["dme"], none, P["dg"], P["dmu"], none, sizes, cnt, B1,
                       B2, ER, s)
    abi.forward_inplace_mt(dtype_id, P["g"], P["mu"], P["nu"], sizes, cnt, B1, B2, EPS, ER, s)
    if True:   # whole-step kernels: all three dtype families
        abi.step_forward(dtype_id, P["p"], P["g"], P["mu"], P["nu"], P["p2"], P["mu2"], P["nu2"], P["u"], sizes, cnt, B1,
                         B2, EPS, ER, -1e-3, s)
        abi.step_backward(dtype_id, P["du"], None, P["mu2"], P["nu2"], P["g"], P["dme"], P["dne"], P["dg"], P["dmu"],
                          P["dnu"], sizes, cnt, B1, B2, EPS, ER, -1e-3, s)
        abi.step_backward(dtype_id, P["du"], P["du"], P["mu2"], P["nu2"], P["g"], none, P["dne"], P["dg"], none,
                          P["dnu"], sizes, cnt, B1, B2, EPS, ER, -1e-3, s)
        # weight decay (L2 / decoupled) + maximize: extra p read, extra dp write
        abi.step_forward(dtype_id, P["p"], P["g"], P["mu"], P["nu"], P["p2"], P["mu2"], P["nu2"], None, sizes, cnt, B1,
                         B2, EPS, ER, -1e-3, s, wd_l2=1e-2, maximize=True)
        abi.step_backward(dtype_id, P["du"], None, P["mu2"], P["nu2"], P["g"], P["dme"], P["dne"], P["dg"], P["dmu"],
                          P["dnu"], sizes, cnt, B1, B2, EPS, ER, -1e-3, s, p=P["p"], dp_out=P["p2"], wd_l2=1e-2,
                          maximize=True)
        abi.step_backward(dtype_id, P["du"], None, P["mu2"], P["nu2"], P["g"], none, none, P["dg"], P["dmu"],
                          P["dnu"], sizes, cnt, B1, B2, EPS, ER, -1e-3, s, dp_out=P["p2"], wd_decoupled=1e-2)
        abi.step_inplace(dtype_id, P["p"], P["g"], P["mu"], P["nu"], sizes, cnt, B1, B2, EPS, ER, -1e-3, s,
                         wd_decoupled=1e-2)
        abi.step_inplace(dtype_id, P["p"], P["g"], P["mu"], P["nu"], sizes, cnt, B1, B2, EPS, ER, -1e-3, s)
    if dtype_id != abi.BF16_F32:   # the six out-of-place reference operators: fp32 / fp64 only
        for i, n in enumerate(sizes):
            if not n:
                continue
            abi.forward_mu(dtype_id, P["g"][i], P["mu"][i], P["mu2"][i], n, B1, s)
            abi.forward_nu(dtype_id, P["g"][i], P["nu"][i], P["nu2"][i], n, B2, s)
            abi.forward_updates(dtype_id, P["mu2"][i], P["nu2"][i], P["u"][i], n, B1, B2, EPS, ER, 2, s)
            abi.backward_mu(dtype_id, P["dme"][i], P["dg"][i], P["dmu"][i], n, B1, s)
            abi.backward_nu(dtype_id, P["dne"][i], P["g"][i], P["dg"][i], P["dnu"][i], n, B2, s)
            abi.backward_updates(dtype_id, P["du"][i], P["u"][i], P["mu2"][i], P["dmu"][i], P["dnu"][i], n, B1, B2, ER, 2, s)
    torch.cuda.synchronize()
    for v in A.values():
        for a in v:
            a.check()


def main():
    sizes = [0, 1, 3, 7, 8, 9, 31, 33, 255, 1023, 1025, 4095, 4096, 4097, 8191, 8193, 12_289, 40_961]
    many = [1 + (i * 37) % 700 for i in range(300)]
    for off in (0, 1, 3):          # 0: vector path; 1, 3: unaligned bases -> element-wise path
        run(abi.F32, torch.float32, torch.float32, sizes, off)
        run(abi.F64, torch.float64, torch.float64, sizes, off)
        run(abi.BF16_F32, torch.float32, torch.bfloat16, sizes, off)
    run(abi.F32, torch.float32, torch.float32, many, 0)     # > 256 leaves: several tables
    print("sanitize_check ok:", abi.kernel_launches(), "launches")


if __name__ == "__main__":
    main()
