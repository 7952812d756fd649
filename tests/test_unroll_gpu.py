"""GPU parity of the autograd dispatch: the fused single-node graph (AdamOp.FusedOp, one launch each way) vs
(a) the reference's three-node graph run on the same kernels and (b) the golden 5-step unroll recorded from the
reference's own AdamOp + compiled adam_op on CPU (tests/golden/make_golden.py)."""
from __future__ import annotations

import pytest

from conftest import assert_bits_equal

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

B1, B2, EPS = 0.9, 0.999, 1e-8


def run_unroll(AdamOp, z, dt, *, mrg, er, fused, multi):
    """Same computation as tests/golden/make_golden.py:run_unroll, on cuda."""
    op = AdamOp(b1=B1, b2=B2, eps=EPS, eps_root=er, inplace=False, fused=fused)
    dev = "cuda"
    p = torch.from_numpy(z["p0"].copy()).to(dev).requires_grad_(True)
    wt = torch.from_numpy(z["w"].copy()).to(dev).requires_grad_(True)
    mu = torch.from_numpy(z["mu0"].copy()).to(dev).requires_grad_(mrg)
    nu = torch.from_numpy(z["nu0"].copy()).to(dev).requires_grad_(mrg)
    t = torch.from_numpy(z["tgt"].copy()).to(dev)
    lr = float(z["lr"])
    leaves = [p, wt] + ([mu, nu] if mrg else [])
    cur, cmu, cnu = p, mu, nu
    for k in range(z["w"].shape[0]):
        g = wt[k] * cur
        if multi:
            (cmu,), (cnu,), (u,) = op.multi([cmu], [cnu], [g], [k + 1])
        else:
            cmu, cnu, u = op(cmu, cnu, g, k + 1)
        cur = cur + (-lr) * u
    outer = ((cur - t) ** 2).sum() + (cmu * t).sum() + (cnu * t * t).sum()
    grads = torch.autograd.grad(outer, leaves)
    torch.cuda.synchronize()
    res = {"p_final": cur, "mu_final": cmu, "nu_final": cnu, "d_p0": grads[0], "d_w": grads[1]}
    if mrg:
        res["d_mu0"], res["d_nu0"] = grads[2], grads[3]
    return {k: v.detach().cpu().numpy() for k, v in res.items()}


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("mrg", [False, True])
@pytest.mark.parametrize("er", [0.0, 1e-8])
def test_unroll_matches_reference_golden(golden, tag, mrg, er):
    from torchopt_b200 import AdamOp
    z = golden(f"adam_unroll_{tag}.npz")
    key = f"mrg{int(mrg)}.er{er:g}."
    fused = run_unroll(AdamOp, z, tag, mrg=mrg, er=er, fused=True, multi=False)
    multi = run_unroll(AdamOp, z, tag, mrg=mrg, er=er, fused=True, multi=True)
    split = run_unroll(AdamOp, z, tag, mrg=mrg, er=er, fused=False, multi=False)
    for name, got in fused.items():
        assert_bits_equal(got, split[name], f"fused vs three-node graph: {name}")
        assert_bits_equal(got, multi[name], f"single-leaf vs multi entry: {name}")
        assert_bits_equal(got, z[key + name], f"vs reference golden: {name}")


def _mlp(seed=0):
    torch.manual_seed(seed)
    net = torch.nn.Sequential(torch.nn.Linear(20, 33), torch.nn.Tanh(), torch.nn.Linear(33, 7)).cuda()
    return net


@pytest.mark.parametrize("mrg", [True, False])
def test_meta_adam_fused_equals_three_node_graph(mrg):
    """MetaAdam over an MLP, 4 inner steps + meta-gradient: the one-launch multi-tensor path and the reference-style
    per-leaf three-node path give bit-identical parameters and meta-gradients; launches drop from 3 per leaf per
    step to 1 per step."""
    import torchopt_b200 as tb
    from torchopt_b200 import optim, transform
    x = torch.randn(16, 20, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    y = torch.randn(16, 7, device="cuda", generator=torch.Generator("cuda").manual_seed(2))
    out = {}
    for fused in (True, False):
        transform.FUSED_GRAPH = fused
        optim.FUSE_STEP = False   # this test is about the chained path (Adam op + torch mul/add)
        try:
            net = _mlp()
            init = [p for p in net.parameters()]
            opt = tb.MetaAdam(net, lr=1e-2, moment_requires_grad=mrg, eps_root=1e-8)
            before = tb.abi.kernel_launches()
            for _ in range(4):
                opt.step(((net(x) - y) ** 2).mean())
            fwd_launches = tb.abi.kernel_launches() - before
            outer = ((net(x) - y) ** 2).mean()
            meta = torch.autograd.grad(outer, init)
            torch.cuda.synchronize()
            out[fused] = ([p.detach().cpu().numpy() for p in opt.params()], [m.cpu().numpy() for m in meta],
                          fwd_launches)
        finally:
            transform.FUSED_GRAPH = True
            optim.FUSE_STEP = True
    assert out[True][2] == 4 and out[False][2] == 4 * 4 * 3
    for a, b in zip(out[True][0], out[False][0]):
        assert_bits_equal(a, b, "unrolled parameters")
    for a, b in zip(out[True][1], out[False][1]):
        assert_bits_equal(a, b, "meta-gradients")


def test_count_host_mirror_and_none_grads():
    """Leaves without a gradient pass through untouched and keep their count; counters are device tensors with the
    reference's layout (0-dim int64) and never force a sync."""
    import torchopt_b200 as tb
    params = [torch.randn(5, device="cuda"), torch.randn(7, device="cuda"), torch.randn(3, device="cuda")]
    tx = tb.scale_by_accelerated_adam.flat(eps_root=1e-8)
    state = tx.init(params)
    assert all(c.dim() == 0 and c.dtype == torch.int64 and c.is_cuda for c in state.count)
    grads = [torch.randn(5, device="cuda"), None, torch.randn(3, device="cuda")]
    upd, state2 = tx.update(grads, state, inplace=False)
    assert upd[1] is None and state2.mu[1] is state.mu[1]
    assert [int(c) for c in state2.count] == [1, 0, 1]
    assert [c._b200_host_count for c in state2.count] == [1, 0, 1]
    # a state whose counters carry no host mirror (e.g. user-built) still works
    bare = tb.ScaleByAdamState(mu=state2.mu, nu=state2.nu, count=[torch.tensor(4, device="cuda") for _ in params])
    upd3, state3 = tx.update(grads, bare, inplace=False)
    assert [int(c) for c in state3.count] == [5, 4, 5]


def test_inplace_adam_matches_torch_optim():
    """torchopt_b200.Adam (in-place forward_ multi-tensor) tracks torch.optim.Adam in fp64, like the reference's
    tests/test_optim.py:352-416 (CUDA, fp64, 25x default tolerances on parameter deltas)."""
    import torchopt_b200 as tb
    torch.manual_seed(0)
    net_a = torch.nn.Sequential(torch.nn.Linear(12, 9), torch.nn.ReLU(), torch.nn.Linear(9, 3)).double().cuda()
    net_b = torch.nn.Sequential(torch.nn.Linear(12, 9), torch.nn.ReLU(), torch.nn.Linear(9, 3)).double().cuda()
    net_b.load_state_dict(net_a.state_dict())
    init = [p.detach().clone() for p in net_a.parameters()]
    oa = tb.Adam(net_a.parameters(), lr=1e-3, weight_decay=1e-2)
    ob = torch.optim.Adam(net_b.parameters(), lr=1e-3, weight_decay=1e-2)
    for i in range(5):
        x = torch.randn(8, 12, dtype=torch.float64, device="cuda")
        for net, o in ((net_a, oa), (net_b, ob)):
            o.zero_grad()
            net(x).pow(2).mean().backward()
            o.step()
    for p0, pa, pb in zip(init, net_a.parameters(), net_b.parameters()):
        torch.testing.assert_close(pa - p0, pb - p0, rtol=2.5e-6, atol=2.5e-6)


def test_native_surface_on_gpu():
    """The seven reference functions through the torch shim: list returns, fresh outputs, aliases for forward_,
    a 0-dim tensor accepted as count, 'Not implemented' for CPU/meta, dtype errors."""
    import torchopt_b200 as tb
    ao = tb.adam_op
    g, mu, nu = (torch.rand(1000, device="cuda") for _ in range(3))
    out = ao.forward_mu(g, mu, 0.9)
    assert out.data_ptr() != mu.data_ptr() and out.shape == mu.shape
    r = ao.backward_mu(g, g, mu, 0.9)
    assert isinstance(r, list) and len(r) == 2
    cnt = torch.tensor(3, device="cuda")
    u1 = ao.forward_updates(mu, nu, 0.9, 0.999, 1e-8, 0.0, cnt)
    u2 = ao.forward_updates(mu, nu, 0.9, 0.999, 1e-8, 0.0, 3)
    assert torch.equal(u1, u2)
    r = ao.forward_(g, mu, nu, 0.9, 0.999, 1e-8, 0.0, 1)
    assert r[0].data_ptr() == g.data_ptr() and r[1].data_ptr() == mu.data_ptr() and r[2].data_ptr() == nu.data_ptr()
    with pytest.raises(RuntimeError, match="Not implemented"):
        ao.forward_mu(g.cpu(), mu.cpu(), 0.9)
    with pytest.raises(RuntimeError, match="not implemented for 'Int'"):
        ao.forward_mu(g.int(), mu.int(), 0.9)
    assert tb.accelerated_op_available() and tb.accelerated_op_available("cuda:0")
    assert not tb.accelerated_op_available("cpu") and not tb.accelerated_op_available("meta")
    # transposed donors: outputs follow the donor's strides, pairing is by storage order (reference behaviour)
    a = torch.rand(8, 6, device="cuda").t()
    o = ao.forward_mu(a, a, 0.5)
    assert o.stride() == a.stride()
