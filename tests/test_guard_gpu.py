"""Out-of-bounds check with guard sentinels around every buffer (compute-sanitizer is closed on this GPU pool):
every kernel family over ragged tails, unaligned bases, NULL patterns and > 256-leaf tables must leave the
guard elements on both sides of every array untouched.  Driver: tools/sanitize_check.py."""
from __future__ import annotations

import importlib.util
import os

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_no_kernel_writes_outside_its_ranges():
    pytest.importorskip("torch")
    spec = importlib.util.spec_from_file_location("sanitize_check", os.path.join(ROOT, "tools", "sanitize_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()


def test_no_kernel_reads_or_writes_outside_its_ranges_guard_pages():
    """tools/guard_pages.cu: every array of every call sits in its own granule of CUDA virtual memory between two UNMAPPED
    granules (flush with the end, then with the start), so an out-of-range LOAD faults too -- what the sentinel test above
    cannot see.  Runs as a separate process: a fault would poison the CUDA context.  The program never provokes a fault
    itself; it ends by asking the driver (pointer attributes) that the bytes just outside every granule are unmapped and
    the bytes just inside are mapped, so a pass is not vacuous."""
    import json
    import subprocess
    exe = os.path.join(ROOT, "build", "guard_pages")
    if not os.path.isfile(exe):
        pytest.fail("build/guard_pages missing: run __graft_entry__.build()")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    rec = json.loads(res.stdout.strip().splitlines()[-1])
    assert rec["guard_pages"] == "ok" and rec["calls"] > 2000 and rec["neighbours_unmapped"] is True
