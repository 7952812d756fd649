"""CUDA-graph capture of whole unrolls (torchopt_b200.graphs): a replay must reproduce the eager computation --
bit-for-bit where only element-wise kernels are involved -- on fresh inputs, with no host launches of the Adam kernels."""
from __future__ import annotations

import pytest
import torch

gpu = pytest.mark.gpu


def test_capture_needs_cuda_when_absent():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import torchopt_b200 as tb
    with pytest.raises(RuntimeError, match="CUDA"):
        tb.graphs.capture(lambda: None)


@gpu
@pytest.mark.parametrize("mrg", [False, True])
def test_captured_unroll_equals_eager_bitwise(mrg):
    """5-step FuncAdam unroll + meta-backward over ragged leaves: replay on NEW inputs == eager on the same inputs."""
    import torchopt_b200 as tb
    dev = torch.device("cuda")
    torch.manual_seed(11)
    sizes = [5, 64, 320, 4097, 36_864, 100_003]
    params = [(torch.randn(k, device=dev) * 0.1).requires_grad_(True) for k in sizes]

    def meta_grad(*ws):
        opt = tb.FuncAdam(1e-2, eps_root=1e-8, moment_requires_grad=mrg)
        cur = list(params)
        for _ in range(5):
            cur = opt.apply_gradients([w * p for w, p in zip(ws, cur)], cur)
        outer = torch.stack([(p * p).sum() for p in cur]).sum()
        return torch.autograd.grad(outer, params)

    ws0 = [torch.rand(k, device=dev) + 0.5 for k in sizes]
    step = tb.graphs.capture(meta_grad, *ws0)
    for trial in range(3):
        ws = [torch.rand(k, device=dev) + 0.25 * (trial + 1) for k in sizes]
        want = meta_grad(*ws)
        before = tb.abi.kernel_launches()
        got = step(*ws)
        torch.cuda.synchronize()
        assert tb.abi.kernel_launches() == before, "a replay must not go through the host launch path"
        for a, b in zip(got, want):
            assert torch.equal(a, b)
    # parameters edited in place are picked up too (they are read by the recorded kernels, not copied)
    with torch.no_grad():
        for p in params:
            p.mul_(1.5)
    want = meta_grad(*ws0)
    got = step(*ws0)
    for a, b in zip(got, want):
        assert torch.equal(a, b)


@gpu
def test_captured_maml_task_accumulates_into_bucket():
    """MetaAdam inner loop on a module + query loss + backward() into a GradBucket, captured once and replayed per task:
    same meta-gradient as the eager loop (matmul-based net; tolerance covers cuBLAS algorithm choice only)."""
    import torch.nn.functional as F
    from torch import nn

    import torchopt_b200 as tb
    from torchopt_b200.parallel import GradBucket
    dev = torch.device("cuda")
    torch.manual_seed(3)
    torch.backends.cuda.matmul.allow_tf32 = False
    net = nn.Sequential(nn.Linear(20, 33), nn.Tanh(), nn.Linear(33, 5)).to(dev)
    slots = [(m, k) for m in net.modules() for k, p in m._parameters.items() if p is not None]
    meta = [m._parameters[k] for m, k in slots]
    bucket = GradBucket(meta)
    tasks = 4
    xs, ys = torch.randn(tasks, 16, 20, device=dev), torch.randint(0, 5, (tasks, 16), device=dev)
    xq, yq = torch.randn(tasks, 24, 20, device=dev), torch.randint(0, 5, (tasks, 24), device=dev)

    def rebind():
        for (m, k), p in zip(slots, meta):
            m._parameters[k] = p

    def task_backward(x_s, y_s, x_q, y_q):
        rebind()
        inner = tb.MetaAdam(net, lr=1e-2, moment_requires_grad=True)
        for _ in range(3):
            inner.step(F.cross_entropy(net(x_s), y_s))
        loss = F.cross_entropy(net(x_q), y_q) / tasks
        rebind()
        loss.backward()
        return loss.detach()

    bucket.zero()
    eager_losses = [task_backward(xs[t], ys[t], xq[t], yq[t]).clone() for t in range(tasks)]
    want = [p.grad.clone() for p in meta]

    step = tb.graphs.capture(task_backward, xs[0], ys[0], xq[0], yq[0], before_each=bucket.zero)
    for _ in range(2):  # two meta-iterations: zero between them, replays accumulate
        bucket.zero()
        losses = [step(xs[t], ys[t], xq[t], yq[t]).clone() for t in range(tasks)]
        torch.cuda.synchronize()
        for a, b in zip(losses, eager_losses):
            torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-7)
        for p, w in zip(meta, want):
            torch.testing.assert_close(p.grad, w, rtol=1e-5, atol=1e-7)
    assert step.replays == 2 * tasks


    # two more instances with their own gradient buffers, replayed side by side on their own streams
    buckets = [GradBucket(meta) for _ in range(2)]
    inst = [tb.graphs.capture(task_backward, xs[0], ys[0], xq[0], yq[0], before_each=b.zero) for b in buckets]
    for b in buckets:
        b.zero()
    issued = tb.graphs.replay_concurrently(inst, ((xs[t], ys[t], xq[t], yq[t]) for t in range(tasks)))
    assert issued == tasks and [i.replays for i in inst] == [2, 2]
    total = [a + b for a, b in zip(buckets[0].views, buckets[1].views)]
    torch.cuda.synchronize()
    for got, w in zip(total, want):
        torch.testing.assert_close(got, w, rtol=1e-5, atol=1e-7)


@gpu
@pytest.mark.parametrize("path", ["whole_step", "chained_dict"])
def test_capture_after_eager_runs_and_no_garbage(path):
    """Eager runs on the default stream followed by a capture of the same function: nothing of the eager graphs may
    survive (a leaf's gradient-accumulator node created on the default stream would make the capture illegal), and a
    step must not leave reference cycles behind."""
    import gc
    import weakref

    import torchopt_b200 as tb
    dev = torch.device("cuda")
    torch.manual_seed(2)
    sizes = [9408, 64, 64, 36_864]
    params = [(torch.randn(k, device=dev) * 0.1).requires_grad_(True) for k in sizes]
    lay = tb.FlatLayout(params)
    w_flat = lay.pack([torch.rand(k, device=dev) + 0.5 for k in sizes])
    probe = []

    def fn():
        cur = lay.pack(params)
        if path == "whole_step":
            opt = tb.FuncAdam(1e-3, eps_root=1e-8)
            (cur,) = opt.apply_gradients([w_flat * cur], [cur])
        else:
            opt = tb.adam(1e-3, eps_root=1e-8)
            tree = {"w": cur}
            state = opt.init(tree)
            updates, state = opt.update({"w": w_flat * cur}, state, params=tree, inplace=False)
            cur = tb.apply_updates(tree, updates, inplace=False)["w"]
        probe[:] = [weakref.ref(cur)]
        return torch.autograd.grad(cur.sum(), params)

    gc.collect()
    gc.disable()
    try:
        want = [g.clone() for g in fn()]
        assert probe[0]() is None, "the unrolled tensors of an eager run outlived it (reference cycle)"
        kept = fn()                                   # results of an eager run kept alive across the capture
        step = tb.graphs.capture(fn)
    finally:
        gc.enable()
    got = step()
    torch.cuda.synchronize()
    for a, b, c in zip(got, want, kept):
        assert torch.equal(a, b) and torch.equal(a, c)


@gpu
def test_persistent_state_inside_a_capture_is_refused():
    """Bias corrections are host scalars: a recorded step would replay with the step count of the recording.  Stepping a
    state that was created outside the capture must therefore fail loudly instead of silently freezing the count."""
    import torchopt_b200 as tb
    dev = torch.device("cuda")
    p = [torch.randn(1000, device=dev).requires_grad_(True)]
    opt = tb.Adam(p, lr=1e-3)                     # persistent state, created eagerly
    p[0].grad = torch.randn(1000, device=dev)
    opt.step()                                     # eager steps are fine
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with pytest.raises(RuntimeError, match="OUTSIDE a CUDA-graph capture"):
        with torch.cuda.graph(graph, stream=side):
            opt.step()
    torch.cuda.synchronize()
    flat = tb.FlatAdam([torch.randn(64, device=dev).requires_grad_(True)], lr=1e-3)
    graph2 = torch.cuda.CUDAGraph()
    with pytest.raises(RuntimeError, match="OUTSIDE a CUDA-graph capture"):
        with torch.cuda.graph(graph2, stream=side):
            flat.step()
    torch.cuda.synchronize()
    opt.step()                                     # and the optimizer still works eagerly afterwards


@gpu
def test_capture_with_leaves_at_different_step_counts():
    """A leaf that gets no gradient in the first inner step has a different step count afterwards; the counters are
    then built by fills (not by a host->device copy, which must not be recorded) and the replay matches eager."""
    import torchopt_b200 as tb
    dev = torch.device("cuda")
    torch.manual_seed(9)
    sizes = [100, 2000, 33]
    params = [(torch.randn(k, device=dev) * 0.2).requires_grad_(True) for k in sizes]

    def fn(w):
        opt = tb.FuncAdam(1e-2, eps_root=1e-8)
        cur = list(params)
        for k in range(3):
            grads = [w[: p.numel()] * p for p in cur]
            if k == 0:
                grads[1] = None
            cur = opt.apply_gradients(grads, cur)
        st = next(s for s in opt.state if isinstance(s, tb.ScaleByAdamState))
        assert [int(c._b200_host_count) for c in st.count] == [3, 2, 3]
        return torch.autograd.grad(torch.stack([(p * p).sum() for p in cur]).sum(), params)

    w0 = torch.rand(2000, device=dev) + 0.5
    step = tb.graphs.capture(fn, w0)
    w1 = torch.rand(2000, device=dev) + 0.25
    want = fn(w1)
    got = step(w1)
    torch.cuda.synchronize()
    for a, b in zip(got, want):
        assert torch.equal(a, b)
