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
