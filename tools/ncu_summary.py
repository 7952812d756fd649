#!/usr/bin/env python
"""Turn the scratch ncu outputs in gpurun_out/ into the committed summaries under profiles/.

    python tools/ncu_summary.py --round r01 --launches gpurun_out/launches.csv --report gpurun_out/prof_r01.ncu-rep
"""
from __future__ import annotations

import argparse
import collections
import csv
import json
import os
import re
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROFILES = os.path.join(ROOT, "profiles")

RAW_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def short_name(name: str) -> str:
    m = re.search(r"stream_kernel<(?:b200adam::)?(\w+)<([^>]*)>,\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+)", name)
    if m:
        return f"stream_kernel<{m.group(1)}<{m.group(2)}>, P={m.group(3)}, U={m.group(4)}, BLOCK={m.group(5)}, MAXL={m.group(6)}>"
    m = re.search(r"stream_kernel<(?:b200adam::)?(\w+)<([^>]*)>", name)
    if m:
        return f"stream_kernel<{m.group(1)}<{m.group(2)}>>"
    return re.sub(r"<.*", "", name)[:70]


def launches(path: str, tag: str, command: str) -> None:
    lines = [l for l in open(path) if l.startswith('"')]
    agg: dict = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        key = (short_name(row["Kernel Name"]), row["Grid Size"], row["Block Size"])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += float(row["Metric Value"])
        n += 1
    total = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if "stream_kernel" in k[0])
    out = [f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none)", "",
           f"Command: `{command}`", "",
           f"{n} launches captured, {total / 1e6:.3f} ms of device time; this repo's kernels (`stream_kernel<...>`): "
           f"{100 * ours / total:.1f} % of it (the rest is torch's synthetic-input generation: randn/rand/mul).", "",
           "Per-launch times are cold-cache and serialised by ncu: compare SHARES, not absolutes.", "",
           "| launches | total ms | share | avg us | grid | block | kernel |", "|---:|---:|---:|---:|---|---|---|"]
    for (name, grid, block), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / total:.2f} % | {a[1] / a[0] / 1e3:.1f} | {grid} | {block} | `{name}` |")
    with open(os.path.join(PROFILES, f"{tag}_launches_summary.md"), "w") as f:
        f.write("\n".join(out) + "\n")
    shutil.copy(path, os.path.join(PROFILES, f"{tag}_launches.csv"))


def report(path: str, tag: str, command: str, n_elems: int, suffix: str = "") -> None:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    with open(os.path.join(PROFILES, f"{tag}_ncu_full{suffix}_raw.csv"), "w") as f:
        f.write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = [f"# {tag}: ncu --set full of the two {'whole-step' if suffix else 'fused'} kernels", "", f"Command: `{command}`", "",
           f"n = {n_elems} fp32 elements per array (arrays of 1 GiB >> 126 MB L2).", ""]
    traffic = {"n_elems": n_elems, "source": f"profiles/{tag}_ncu_full{suffix}_raw.csv"}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        is_bwd = "Bwd" in name
        per_elem = 36 if is_bwd else (28 if "StepFwd" in name else 24)
        alg = per_elem * n_elems
        rd = float(r[hdr.index("dram__bytes_read.sum")])
        wr = float(r[hdr.index("dram__bytes_write.sum")])
        u_rd, u_wr = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
        tot = rd * scale[u_rd] + wr * scale[u_wr]
        key = ("step_" if "Step" in name else "") + ("backward" if is_bwd else "forward")
        traffic[f"{key}_dram_bytes_per_launch"] = tot
        traffic[f"{key}_dram_bytes_per_elem"] = tot / n_elems
        out += [f"## `{short_name(name)}`", "",
                f"grid {r[hdr.index('Grid Size')]}, block {r[hdr.index('Block Size')]}; algorithmic bytes "
                f"{alg / 1e9:.4f} GB ({per_elem} B/elem); measured DRAM traffic {tot / 1e9:.4f} GB "
                f"({tot / alg:.3f} x algorithmic).", "", "| metric | value | unit |", "|---|---:|---|"]
        for m in RAW_METRICS:
            if m in hdr:
                i = hdr.index(m)
                out.append(f"| `{m}` | {r[i]} | {units[i]} |")
        out.append("")
    with open(os.path.join(PROFILES, f"{tag}_ncu_full{suffix}_summary.md"), "w") as f:
        f.write("\n".join(out) + "\n")
    tpath = os.path.join(PROFILES, "ncu_traffic.json")
    merged = {}
    if os.path.isfile(tpath):
        merged = json.load(open(tpath))
    if any(k.startswith("step_") for k in traffic):   # a capture of the step kernels only adds its own keys
        merged.update({k: v for k, v in traffic.items() if k.startswith("step_")})
        merged["step_source"] = traffic["source"]
    else:
        merged.update(traffic)
    with open(tpath, "w") as f:
        json.dump(merged, f, indent=1)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r01")
    ap.add_argument("--launches")
    ap.add_argument("--launches-cmd", default="")
    ap.add_argument("--report")
    ap.add_argument("--report-cmd", default="")
    ap.add_argument("--report-n", type=int, default=1 << 28)
    ap.add_argument("--suffix", default="", help="e.g. _step for a capture of the whole-step kernels")
    a = ap.parse_args()
    os.makedirs(PROFILES, exist_ok=True)
    if a.launches:
        launches(a.launches, a.round, a.launches_cmd)
    if a.report:
        report(a.report, a.round, a.report_cmd, a.report_n, a.suffix)
