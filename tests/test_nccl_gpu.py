"""Task-parallel MAML over NCCL on two real GPUs (BASELINE configs[3]): the meta-gradient summed by
ONE all-reduce of the flat bucket equals the single-process gradient, and every rank ends with identical meta-parameters.
Skipped on a one-GPU box; the gloo world-size-2 tests in tests/test_host_logic.py cover the same logic on CPU."""
from __future__ import annotations

import json
import os
import socket
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "cuda_graph"])
def test_two_rank_nccl_meta_gradient_equals_single_process(graph):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "examples", "maml_task_parallel.py"),
           "--tasks", "8", "--iters", "2", "--check"] + (["--graph", "--streams", "4"] if graph else [])
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    rec = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    assert rec["n_gpus"] == 2 and rec["tasks_per_rank"] == 4
    assert rec["max_rel_diff_vs_single_process"] < 1e-4     # --check also asserts identical meta-parameters on all ranks
