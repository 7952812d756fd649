"""Flat-buffer layout (torchopt_b200.flat): FlatMetaAdam / FlatAdam must compute, element for element,
what MetaAdam / Adam compute on the per-leaf layout -- with ONE launch over ONE range per step each way."""
from __future__ import annotations

import pytest
import torch
from torch import nn

gpu = pytest.mark.gpu


def test_layout_offsets_pack_and_views_cpu():
    from torchopt_b200.flat import FlatLayout
    leaves = [torch.arange(k, dtype=torch.float32).reshape(shape) + 1 for k, shape in
              ((5, (5,)), (64, (8, 8)), (33, (3, 11)), (1, ()), (32, (32,)))]
    lay = FlatLayout(leaves)
    assert lay.offsets == [0, 32, 96, 160, 192] and lay.total == 224      # 128-byte (32 fp32) aligned starts
    flat = lay.pack(leaves)
    assert flat.numel() == lay.total
    for v, p, o in zip(lay.views(flat), leaves, lay.offsets):
        assert v.shape == p.shape and torch.equal(v, p) and v.data_ptr() == flat.data_ptr() + 4 * o
    mask = torch.ones(lay.total, dtype=torch.bool)
    for o, k in zip(lay.offsets, lay.numels):
        mask[o:o + k] = False
    assert torch.all(flat[mask] == 0), "padding holds zeros"
    # gradient of a function of the views w.r.t. the flat buffer: one tensor, zeros in the padding and for unused views
    flat = flat.clone().requires_grad_(True)
    vs = lay.views(flat)
    (g,) = torch.autograd.grad((vs[1] * 2).sum() + (vs[3] * 3).sum(), [flat])
    want = torch.zeros(lay.total)
    want[32:96], want[160] = 2.0, 3.0
    assert torch.equal(g, want)
    with pytest.raises(ValueError):
        FlatLayout([torch.zeros(2), torch.zeros(2, dtype=torch.float64)])


def test_padded_flat_layouts_refuse_zero_epsilons_cpu():
    """eps == eps_root == 0 would turn the zero padding between leaves into 0/0 = NaN: refused up front."""
    import torchopt_b200 as tb
    net = nn.Linear(3, 2)
    with pytest.raises(ValueError, match="padding"):
        tb.FlatMetaAdam(net, lr=1e-3, eps=0.0, eps_root=0.0)
    with pytest.raises(ValueError, match="padding"):
        tb.FlatAdam(net.parameters(), lr=1e-3, eps=0.0)


class Elementwise(nn.Module):
    """A 'network' made of element-wise maps only, so that the two layouts must agree bit for bit."""

    def __init__(self, sizes, dev, dtype=torch.float32):
        super().__init__()
        self.ps = nn.ParameterList([nn.Parameter((torch.randn(k, device=dev) * 0.3).to(dtype)) for k in sizes])

    def forward(self, ws):
        return torch.stack([(w * p * p).sum() for w, p in zip(ws, self.ps)]).sum()


SIZES = [5, 64, 321, 4097, 36_864]


@gpu
@pytest.mark.parametrize("variant,dtype", [("adam", torch.float32), ("adam_l2", torch.float32), ("adamw", torch.float32),
                                           ("adam", torch.float64), ("adamw", torch.bfloat16)])
def test_flat_meta_adam_equals_per_leaf(variant, dtype):
    import torchopt_b200 as tb
    dev = torch.device("cuda")
    torch.manual_seed(5)
    net = Elementwise(SIZES, dev, dtype)
    ws = [[(torch.rand(k, device=dev) + 0.5).to(dtype) for k in SIZES] for _ in range(4)]
    kw = {"adam": {}, "adam_l2": {"weight_decay": 0.05}, "adamw": {"weight_decay": 0.05}}[variant]
    per_leaf = tb.MetaAdamW if variant == "adamw" else tb.MetaAdam
    flat_cls = tb.FlatMetaAdamW if variant == "adamw" else tb.FlatMetaAdam
    meta = list(net.ps)
    slots = [(m, k) for m in net.modules() for k, p in m._parameters.items() if p is not None]

    def run(cls):
        for (m, k), p in zip(slots, meta):
            m._parameters[k] = p
        inner = cls(net, lr=1e-2, eps_root=1e-8, moment_requires_grad=True, **kw)
        l0 = tb.abi.kernel_launches()
        for k in range(3):
            inner.step(net(ws[k]))
        final = [p.detach().clone() for p in (m._parameters[k] for m, k in slots)]
        grads = torch.autograd.grad(net(ws[3]), meta)
        launches = tb.abi.kernel_launches() - l0
        for (m, k), p in zip(slots, meta):
            m._parameters[k] = p
        return final, grads, launches

    p_leaf, g_leaf, n_leaf = run(per_leaf)
    p_flat, g_flat, n_flat = run(flat_cls)
    assert n_leaf == n_flat == 6, "3 forward + 3 backward launches in either layout"
    for a, b in zip(p_flat, p_leaf):
        assert torch.equal(a, b), "unrolled parameters are bit-identical in both layouts"
    # meta-gradients: the same addends, but a parameter used several times by the loss has >2 gradient contributions and
    # the autograd engine sums them in a different order (per view first, then with the step's dp) -> last-bit noise
    rtol, atol = {torch.float32: (2e-5, 1e-7), torch.float64: (1e-12, 1e-15), torch.bfloat16: (5e-2, 1e-2)}[dtype]
    for a, b in zip(g_flat, g_leaf):
        torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


@gpu
def test_flat_meta_adam_on_a_real_module_and_state_swap():
    import torch.nn.functional as F

    import torchopt_b200 as tb
    dev = torch.device("cuda")
    torch.manual_seed(6)
    torch.backends.cuda.matmul.allow_tf32 = False
    net = nn.Sequential(nn.Linear(20, 33), nn.Tanh(), nn.Linear(33, 5)).to(dev)
    x, y = torch.randn(16, 20, device=dev), torch.randint(0, 5, (16,), device=dev)
    meta = [p for p in net.parameters()]
    slots = [(m, k) for m in net.modules() for k, p in m._parameters.items() if p is not None]

    def run(cls):
        inner = cls(net, lr=1e-2)
        for _ in range(3):
            inner.step(F.cross_entropy(net(x), y))
        grads = torch.autograd.grad(F.cross_entropy(net(x), y), meta)
        for (m, k), p in zip(slots, meta):
            m._parameters[k] = p
        return grads

    for a, b in zip(run(tb.FlatMetaAdam), run(tb.MetaAdam)):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-7)

    # state save / restore is a swap of references: stepping again from a restored state reproduces the same step
    inner = tb.FlatMetaAdam(net, lr=1e-2)
    inner.step(F.cross_entropy(net(x), y))
    saved = inner.state_dict()
    inner.step(F.cross_entropy(net(x), y))
    after = inner.flat_p.detach().clone()
    inner.load_state_dict(saved)
    inner.step(F.cross_entropy(net(x), y))
    assert torch.equal(inner.flat_p.detach(), after)
    inner.restore()
    assert all(a is b for a, b in zip(net.parameters(), meta))


@gpu
@pytest.mark.parametrize("decoupled", [False, True])
def test_flat_adam_inplace_equals_per_leaf_bitwise(decoupled):
    import torchopt_b200 as tb
    dev = torch.device("cuda")
    torch.manual_seed(7)
    net_a = Elementwise(SIZES, dev)
    net_b = Elementwise(SIZES, dev)
    net_b.load_state_dict(net_a.state_dict())
    ws = [[torch.rand(k, device=dev) + 0.5 for k in SIZES] for _ in range(4)]
    kw = {"weight_decay": 0.05, "lr": 1e-2, "eps_root": 1e-8}
    opt_a = (tb.AdamW if decoupled else tb.Adam)(net_a.parameters(), **kw)
    opt_b = (tb.FlatAdamW if decoupled else tb.FlatAdam)(net_b.parameters(), **kw)
    for p, q in zip(net_a.parameters(), net_b.parameters()):
        assert torch.equal(p, q), "re-pointing p.data keeps the values"
    for k in range(4):
        opt_a.zero_grad()
        opt_b.zero_grad()
        net_a(ws[k]).backward()
        net_b(ws[k]).backward()
        assert all(p.grad.data_ptr() == gv.data_ptr() for p, gv in zip(net_b.parameters(), opt_b.g_views)), \
            "backward accumulates straight into the flat gradient buffer"
        if k == 2:   # a foreign gradient tensor is gathered into the flat buffer
            first = next(iter(net_b.parameters()))
            first.grad = first.grad.clone()
        l0 = tb.abi.kernel_launches()
        g_before = opt_b.flat_g.clone()
        opt_b.step()
        assert tb.abi.kernel_launches() - l0 == 1
        assert torch.equal(opt_b.flat_g, g_before), "FlatAdam leaves its gradient buffer untouched (28 B/elem step)"
        opt_a.step()
        for p, q in zip(net_a.parameters(), net_b.parameters()):
            assert torch.equal(p, q)
