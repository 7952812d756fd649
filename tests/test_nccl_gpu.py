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


def test_parameter_list_spanning_two_devices():
    """Leaves on cuda:0 and cuda:1 in ONE optimizer (group by device): the flat-state engine
    forms one group -- one launch, on that device's current stream under a device guard -- per (device, dtype), and the
    results equal running each device's leaves through their own optimizer; same for the functional transformation."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torchopt_b200 as tb
    sizes, devs = [7, 4097, 33, 70_000], ["cuda:0", "cuda:1", "cuda:1", "cuda:0"]
    gens = {d: torch.Generator(d).manual_seed(3) for d in set(devs)}
    p0 = [torch.randn(k, device=d, generator=gens[d]) * 0.3 for k, d in zip(sizes, devs)]
    ws = [torch.rand(k, device=d, generator=gens[d]) + 0.5 for k, d in zip(sizes, devs)]

    def unroll(index):
        params = [p0[i].clone().requires_grad_(True) for i in index]
        opt = tb.FuncAdam(1e-2, eps_root=1e-8, weight_decay=1e-2, moment_requires_grad=True)
        cur = list(params)
        launches0 = tb.abi.kernel_launches()
        for _ in range(3):
            cur = opt.apply_gradients([ws[i] * p for i, p in zip(index, cur)], cur)
        launches = tb.abi.kernel_launches() - launches0
        outer = sum((q * q).sum().to("cuda:0") for q in cur)
        meta = torch.autograd.grad(outer, params)
        for d in set(devs):
            torch.cuda.synchronize(d)
        return [q.detach() for q in cur], meta, launches

    both, meta_both, launches = unroll(range(4))
    assert launches == 2 * 3, "one forward launch per device group per step"
    for dev_index in ([0, 3], [1, 2]):
        alone, meta_alone, _ = unroll(dev_index)
        for j, i in enumerate(dev_index):
            assert both[i].device == p0[i].device
            assert torch.equal(both[i], alone[j]) and torch.equal(meta_both[i], meta_alone[j])
    # the functional transformation (per-leaf state lists, AdamOp.multi) groups by device as well
    opt = tb.adam(1e-2, eps_root=1e-8)
    params = [p.clone().requires_grad_(True) for p in p0]
    state = opt.init(params)
    updates, state = opt.update([w * p for w, p in zip(ws, params)], state, params=params, inplace=False)
    new = tb.apply_updates(params, updates, inplace=False)
    for i in range(4):
        ref_opt = tb.adam(1e-2, eps_root=1e-8)
        st = ref_opt.init([params[i]])
        up, _ = ref_opt.update([ws[i] * params[i]], st, params=[params[i]], inplace=False)
        assert torch.equal(new[i], tb.apply_updates([params[i]], up, inplace=False)[0])
