"""The default state layout of MetaAdam / FuncAdam / Adam: flat moment buffers + host step counts (``FlatAdamState``,
``adam_op.flat_step``) against the per-leaf reference layout (``optim.FLAT_STATE = False``: ScaleByAdamState lists,
4L autograd inputs per step).  Everything -- parameters, moments, counts, meta-gradients, the reference's
skip-leaves-without-gradient semantics -- must be bit-identical; what changes is the host-side bookkeeping only."""
from __future__ import annotations

import contextlib

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@contextlib.contextmanager
def per_leaf_state():
    from torchopt_b200 import optim
    optim.FLAT_STATE = False
    try:
        yield
    finally:
        optim.FLAT_STATE = True


class Net(torch.nn.Module):
    """Small MLP with an unused sub-module and an unused parameter, like the reference's test model
    (tests/helpers.py:95-113): their gradients are None, so every step exercises the skip path."""

    def __init__(self, dtype=torch.float32):
        super().__init__()
        torch.manual_seed(0)
        self.a = torch.nn.Linear(20, 33)
        self.unused_module = torch.nn.Linear(3, 3, bias=False)
        self.b = torch.nn.Linear(33, 7)
        self.unused_parameter = torch.nn.Parameter(torch.zeros(2, 2))
        self.to(dtype)

    def forward(self, x):
        return self.b(torch.tanh(self.a(x)))


def _unused(net):
    """indices (in MetaAdam's slot order == named_parameters order) of the parameters that never get a gradient"""
    return [i for i, (name, _) in enumerate(net.named_parameters()) if "unused" in name]


def _expected_counts(net, steps):
    return [0 if "unused" in name else steps for name, _ in net.named_parameters()]


def _data(dtype=torch.float32):
    x = torch.randn(16, 20, device="cuda", generator=torch.Generator("cuda").manual_seed(1)).to(dtype)
    y = torch.randn(16, 7, device="cuda", generator=torch.Generator("cuda").manual_seed(2)).to(dtype)
    return x, y


def _meta_unroll(cls_name, dtype, mrg, kw, steps=4):
    import torchopt_b200 as tb
    net = Net(dtype).cuda()
    init = list(net.parameters())
    x, y = _data(dtype)
    opt = getattr(tb, cls_name)(net, lr=1e-2, eps_root=1e-8, moment_requires_grad=mrg, **kw)
    before = tb.abi.kernel_launches()
    for _ in range(steps):
        opt.step(((net(x) - y) ** 2).mean())
    launches = tb.abi.kernel_launches() - before
    st = next(s for s in opt.state if isinstance(s, tb.ScaleByAdamState))
    outer = ((net(x) - y) ** 2).mean() + sum((m * m).sum() for m in st.mu) * 0.5
    meta = torch.autograd.grad(outer, init, allow_unused=True)
    torch.cuda.synchronize()
    return ([p.detach() for p in opt.params()], meta, [m.detach() for m in st.mu], [v.detach() for v in st.nu],
            [int(c) for c in st.count], launches, init, _unused(net), _expected_counts(net, steps))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.bfloat16])
@pytest.mark.parametrize("mrg", [True, False])
@pytest.mark.parametrize("cls_name,kw", [("MetaAdam", {}), ("MetaAdam", {"weight_decay": 1e-2, "maximize": True}),
                                         ("MetaAdamW", {"weight_decay": 5e-2})])
def test_meta_adam_flat_state_equals_per_leaf_state(dtype, mrg, cls_name, kw):
    flat = _meta_unroll(cls_name, dtype, mrg, kw)
    with per_leaf_state():
        leaf = _meta_unroll(cls_name, dtype, mrg, kw)
    assert flat[5] == leaf[5] == 4                      # one launch per inner step either way
    assert flat[4] == leaf[4] == flat[8]                # unused leaves (unused_module.weight, unused_parameter) never count
    assert len(flat[7]) == 2
    for name, a, b in zip(("params", "meta", "mu", "nu"), flat[:4], leaf[:4]):
        for x, y in zip(a, b):
            assert (x is None) == (y is None), name
            if x is not None:
                assert x.shape == y.shape and torch.equal(x, y), name
    # skipped leaves: the module keeps the ORIGINAL tensors, their moments stay zero
    for i in flat[7]:
        assert flat[0][i].data_ptr() == flat[6][i].data_ptr() and not flat[2][i].any() and not flat[3][i].any()
        assert flat[1][i] is None


def test_flat_step_autograd_fan_in_and_saved_tensors():
    """The point of the layout: one node per step with 2K+2 inputs (K = leaves with a gradient), not 4L."""
    import torchopt_b200 as tb
    net = Net().cuda()
    x, y = _data()
    opt = tb.MetaAdam(net, lr=1e-2)
    opt.step(((net(x) - y) ** 2).mean())
    node = net.a.weight.grad_fn
    assert node is net.b.bias.grad_fn and node is not None
    assert len(node.next_functions) == 2 * 4 + 2
    assert opt._flat is not None and opt._leaf is None and len(opt._flat.groups) == 1
    assert opt._flat.counts == _expected_counts(net, 1)


def test_state_dict_roundtrip_between_tasks_and_initial_moment_gradients():
    """The reference's MAML loop (examples/few-shot/maml_omniglot.py:136-165): extract the optimizer state once,
    recover it per task.  Reading the state before stepping hands out tensors the next step must consume (gradients
    w.r.t. the INITIAL moments flow to exactly those tensors)."""
    import torchopt_b200 as tb
    results = {}
    for mode in ("flat", "leaf"):
        ctx = per_leaf_state() if mode == "leaf" else contextlib.nullcontext()
        with ctx:
            net = Net().cuda()
            x, y = _data()
            init = list(net.parameters())
            opt = tb.MetaAdam(net, lr=1e-2, moment_requires_grad=True)
            opt.step(((net(x) - y) ** 2).mean())                  # state now exists (count 1)
            tb.stop_gradient(opt), tb.stop_gradient(net)
            net_state, opt_state = tb.extract_state_dict(net), tb.extract_state_dict(opt)
            adam0 = next(s for s in opt_state if isinstance(s, tb.ScaleByAdamState))
            out = []
            for task in range(2):
                xt = x + 0.1 * task
                for _ in range(2):
                    opt.step(((net(xt) - y) ** 2).mean())
                outer = ((net(xt) - y) ** 2).mean()
                used = [i for i in range(6) if i not in _unused(net)]
                wrt = [p for p in (v for d in net_state.params for v in d.values())] + [adam0.mu[used[0]], adam0.nu[used[3]]]
                grads = torch.autograd.grad(outer, wrt, allow_unused=True)
                st = next(s for s in opt.state if isinstance(s, tb.ScaleByAdamState))
                out.append(([g for g in grads], [int(c) for c in st.count], [p.detach() for p in opt.params()]))
                tb.recover_state_dict(net, net_state)
                tb.recover_state_dict(opt, opt_state)
            torch.cuda.synchronize()
            results[mode] = out
            del init
    for a, b in zip(results["flat"], results["leaf"]):
        assert a[1] == b[1] and sorted(a[1]) == [0, 0, 3, 3, 3, 3]
        for x1, x2 in zip(a[0] + a[2], b[0] + b[2]):
            assert (x1 is None) == (x2 is None)
            if x1 is not None:
                assert torch.equal(x1, x2)
        assert a[0][-1] is not None and a[0][-2] is not None     # gradients w.r.t. the extracted initial moments


def test_func_adam_mixed_dtype_groups_and_inplace():
    """fp32 and fp64 leaves in one parameter list form two groups (two launches); inplace=True updates the raw values."""
    import torchopt_b200 as tb
    gen = torch.Generator("cuda").manual_seed(9)
    sizes = [(7,), (130, 3), (), (4097,)]
    dtypes = [torch.float32, torch.float64, torch.float32, torch.float64]

    def run(inplace):
        params = [(torch.randn(s, device="cuda", generator=torch.Generator("cuda").manual_seed(i)) * 0.3).to(dt)
                  .requires_grad_(True) for i, (s, dt) in enumerate(zip(sizes, dtypes))]
        opt = tb.FuncAdam(1e-2, eps_root=1e-8, weight_decay=1e-2)
        cur = list(params)
        l0 = tb.abi.kernel_launches()
        for k in range(3):
            grads = [(p.detach() if inplace else p) * (k + 1.5) for p in cur]
            cur = opt.apply_gradients(grads, cur, inplace=inplace)
        launches = tb.abi.kernel_launches() - l0
        st = next(s for s in opt.state if isinstance(s, tb.ScaleByAdamState))
        torch.cuda.synchronize()
        return [p.detach().clone() for p in cur], [m.detach().clone() for m in st.mu], [int(c) for c in st.count], launches

    for inplace in (False, True):
        a = run(inplace)
        with per_leaf_state():
            b = run(inplace)
        assert a[2] == b[2] == [3, 3, 3, 3] and a[3] == b[3] == 6
        for x, y in zip(a[0] + a[1], b[0] + b[1]):
            assert x.dtype == y.dtype and torch.equal(x, y)
    del gen


def test_inplace_adam_flat_state_equals_per_leaf_state_and_loaded_state():
    import torchopt_b200 as tb

    def run(load_midway):
        net = Net().cuda()
        opt = tb.AdamW(net.parameters(), lr=1e-3, eps_root=1e-8, weight_decay=1e-2)
        gen = torch.Generator("cuda").manual_seed(5)
        for k in range(4):
            x = torch.randn(8, 20, device="cuda", generator=gen)
            opt.zero_grad()
            net(x).pow(2).mean().backward()
            opt.step()
            if load_midway and k == 1:
                opt.load_state_dict(tb.extract_state_dict(opt, by="deepcopy"))
        st = next(s for s in opt.state if isinstance(s, tb.ScaleByAdamState))
        torch.cuda.synchronize()
        return [p.detach().clone() for p in net.parameters()], [m.clone() for m in st.mu], [int(c) for c in st.count]

    a, c = run(False), run(True)
    with per_leaf_state():
        b = run(False)
    assert a[2] == b[2] == c[2] and sorted(a[2]) == [0, 0, 4, 4, 4, 4]
    for x, y, z in zip(a[0] + a[1], b[0] + b[1], c[0] + c[1]):
        assert torch.equal(x, y) and torch.equal(x, z)


def test_persistent_flat_state_refuses_to_be_stepped_inside_a_capture():
    import torchopt_b200 as tb
    p = [torch.randn(64, device="cuda")]
    opt = tb.FuncAdam(1e-3)
    opt.apply_gradients([torch.randn(64, device="cuda")], p)
    g = torch.randn(64, device="cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with pytest.raises(RuntimeError, match="OUTSIDE a CUDA-graph capture"):
        with torch.cuda.graph(graph, stream=side):
            opt.apply_gradients([g], p)
