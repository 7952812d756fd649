#!/usr/bin/env python
"""Opcode histogram of the hot kernels in the shipped libb200adam.so (cuobjdump -sass; no GPU needed).

    python tools/sass_summary.py --round r02      -> profiles/r02_sass_summary.md

What it is evidence for: the library carries sm_100a SASS only; the streaming kernels use 256-bit global accesses
(LDG.E...256 / STG.E...256), keep everything in registers (no LDL / STL spills), evaluate IEEE division / square root
without FMA contraction of the reference's expressions (the FFMA / DFMA present belong to the div / sqrt expansions and
to the explicit fma of the weight-decay terms), carry the reference's fp64 tail in backward_updates (DMUL + F2F), and
implement programmatic dependent launch (ACQBULK = griddepcontrol.wait, PREEXIT = griddepcontrol.launch_dependents).
"""
from __future__ import annotations

import argparse
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "torchopt_b200", "libb200adam.so")

HOT = [
    ("fused forward fp32 (headline)", r"stream_kernelINS_4FwdHIfffLb0EEELi8ELi\d+ELi512ELi1EE"),
    ("fused backward fp32, all gradients present (headline, roofline kernel)", r"stream_kernelINS_4BwdHIfffLi0EEELi8ELi\d+ELi512ELi1EE"),
    ("whole-step forward fp32", r"stream_kernelINS_8StepFwdHIfffLb0EEELi8ELi\d+ELi512ELi1EE"),
    ("whole-step backward fp32, all gradients present", r"stream_kernelINS_8StepBwdHIfffLi0EEELi8ELi\d+ELi\d+ELi1EE"),
    ("fused forward fp64", r"stream_kernelINS_4FwdHIdddLb0EEELi\d+ELi\d+ELi\d+ELi1EE"),
    ("fused backward fp64, all gradients present", r"stream_kernelINS_4BwdHIdddLi0EEELi\d+ELi\d+ELi\d+ELi1EE"),
    ("fused forward bf16-grad / fp32-state", r"stream_kernelINS_4FwdHINS_6bf16_tEffLb0EEELi8ELi\d+ELi512ELi1EE"),
    ("fused backward bf16-grad / fp32-state", r"stream_kernelINS_4BwdHINS_6bf16_tEffLi0EEELi8ELi\d+ELi512ELi1EE"),
]
GROUPS = [("LDG", r"^LDG"), ("STG", r"^STG"), ("LDL/STL (spills)", r"^(LDL|STL)"), ("LDC/ULDC (kernel parameters)", r"^U?LDC"),
          ("FMUL", r"^FMUL"), ("FADD", r"^FADD"), ("FFMA", r"^FFMA"), ("MUFU (rcp/rsq seeds)", r"^MUFU"),
          ("DMUL", r"^DMUL"), ("DADD", r"^DADD"), ("DFMA", r"^DFMA"), ("F2F (fp32<->fp64)", r"^F2F"),
          ("ACQBULK (griddepcontrol.wait)", r"^ACQBULK"), ("PREEXIT (launch_dependents)", r"^PREEXIT"),
          ("UTCMMA / HMMA / TMA (UBLKCP, UTMALDG)", r"^(UTC|HMMA|UBLKCP|UTMA)")]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r02")
    a = ap.parse_args()
    elf = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True, check=True).stdout
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    archs = sorted(set(re.findall(r"sm_\d+a?", elf)))
    funcs, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            cur.append(m.group(1))
    out = [f"# {a.round}: SASS opcode summary of the shipped libb200adam.so", "",
           f"`cuobjdump -lelf`: {', '.join(archs)} only ({len(funcs)} kernels, no PTX).  Produced by `python tools/sass_summary.py` "
           f"on the build box (no GPU needed); the same check runs in `tests/test_abi.py`.", ""]
    for title, pat in HOT:
        names = [n for n in funcs if re.search(pat, n)]
        if not names:
            out += [f"## {title}", "", f"(no kernel matches `{pat}`)", ""]
            continue
        name = names[0]
        ops = funcs[name]
        hist = collections.Counter(ops)
        out += [f"## {title}", "", f"`{name}` -- {len(ops)} instructions", "", "| group | count | opcodes |", "|---|---:|---|"]
        for label, rx in GROUPS:
            sel = {k: v for k, v in hist.items() if re.match(rx, k)}
            detail = ", ".join(f"`{k}` x{v}" for k, v in sorted(sel.items(), key=lambda kv: -kv[1])[:6])
            out.append(f"| {label} | {sum(sel.values())} | {detail} |")
        out.append("")
    with open(os.path.join(ROOT, "profiles", f"{a.round}_sass_summary.md"), "w") as f:
        f.write("\n".join(out) + "\n")
    print("\n".join(out[:40]))


if __name__ == "__main__":
    main()
