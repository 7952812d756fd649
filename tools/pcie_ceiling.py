#!/usr/bin/env python
"""What the host<->device link of this box delivers to plain pinned-memory copies: the ceiling of bench.py's `e2e`
leg (which moves 24 B/elem each way).  H2D alone, D2H alone, and both directions at once on two streams.

    python tools/pcie_ceiling.py [MiB per buffer = 1024]
"""
import json
import sys

import torch

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = mib << 20
dev = torch.device("cuda")
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def h2d():
    d_in.copy_(h_in, non_blocking=True)


def d2h():
    h_out.copy_(d_out, non_blocking=True)


def both():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur), s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
    cur.wait_stream(s1), cur.wait_stream(s2)


gb = n / 1e9
res = {"buffer_MiB": mib, "h2d_GBps": gb / (timed(h2d) * 1e-3), "d2h_GBps": gb / (timed(d2h) * 1e-3)}
ms = timed(both)
res["bidir_each_GBps"] = gb / (ms * 1e-3)
res["e2e_ceiling_gelem_per_s_at_24B_each_way"] = res["bidir_each_GBps"] / 24.0
print(json.dumps(res))

# the e2e pattern without any compute or cross-stream dependency: 6 input + 6 output arrays, copied slice by slice
# (16 MiB per array per slice) on two streams -- separates "link ceiling for this access pattern" from pipeline stalls
m = 1 << 27
sl = 1 << 22
hin = [torch.empty(m, dtype=torch.float32).pin_memory() for _ in range(6)]
hout = [torch.empty(m, dtype=torch.float32).pin_memory() for _ in range(6)]
dbuf = [torch.zeros(sl, dtype=torch.float32, device=dev) for _ in range(12)]


def sliced():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur), s2.wait_stream(cur)
    for off in range(0, m, sl):
        with torch.cuda.stream(s1):
            for a in range(6):
                dbuf[a].copy_(hin[a][off:off + sl], non_blocking=True)
        with torch.cuda.stream(s2):
            for a in range(6):
                hout[a][off:off + sl].copy_(dbuf[6 + a], non_blocking=True)
    cur.wait_stream(s1), cur.wait_stream(s2)


ms = timed(sliced, reps=3)
each = 6 * 4 * m / 1e9 / (ms * 1e-3)
print(json.dumps({"pattern": "6+6 arrays x 2^27 fp32, 16 MiB slices, no dependencies", "bidir_each_GBps": each,
                  "gelem_per_s": m / (ms * 1e-3) / 1e9}))
