"""Leaf layouts the multi-tensor entry points must survive: gradients handed back by autograd can be
expanded (stride 0), permuted or channels_last views, while the kernels walk every array of a leaf linearly.  The shim
re-lays such tensors out like the leaf's donor (``conform`` in torch_shim.cpp); results must equal the same computation
on row-major copies, element for element.  Also: output pooling (``FlatAlloc``) must not let a tiny tensor pin the large
outputs of the same call."""
from __future__ import annotations

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _leaves():
    gen = torch.Generator("cuda").manual_seed(11)
    shapes = [(4, 8, 5, 5), (33, 17), (129,), (2, 3, 4, 6)]
    return [torch.randn(s, device="cuda", generator=gen) * 0.3 for s in shapes]


def _run(params, grads_fn, douts, recipe, transform):
    import torchopt_b200 as tb
    params = [p.detach().clone(memory_format=torch.preserve_format).requires_grad_(True) for p in params]
    if transform:
        opt = tb.adam(1e-2, eps_root=1e-8, moment_requires_grad=True, **recipe)
        cur, state = list(params), opt.init(params)
        for _ in range(2):
            updates, state = opt.update(grads_fn(cur), state, params=cur, inplace=False)
            cur = tb.apply_updates(cur, updates, inplace=False)
    else:
        opt = tb.FuncAdam(1e-2, eps_root=1e-8, moment_requires_grad=True, **recipe)
        cur = list(params)
        for _ in range(2):
            cur = opt.apply_gradients(grads_fn(cur), cur)
    outer = sum((c * d).sum() for c, d in zip(cur, douts))
    meta = torch.autograd.grad(outer, params)
    torch.cuda.synchronize()
    return [c.detach() for c in cur], list(meta)


@pytest.mark.parametrize("transform", [False, True], ids=["whole_step", "scale_by_adam"])
@pytest.mark.parametrize("recipe", [dict(), dict(weight_decay=1e-2)], ids=["plain", "l2"])
def test_expanded_and_permuted_gradients(transform, recipe):
    """stride-0 gradients (d sum(p) / dp) and transposed-view gradients vs their row-major copies."""
    base = _leaves()
    ws = [torch.rand_like(p) + 0.5 for p in base]
    douts = [torch.randn_like(p) for p in base]

    def odd_grads(cur):
        out = []
        for k, (p, w) in enumerate(zip(cur, ws)):
            if k % 2 == 0:   # expanded: every element shares one storage location (what sum().backward() produces)
                out.append((p.sum() * 1e-3 + 0.25).expand(p.shape))
            else:            # a transposed view: same logical values as w * p, column-major storage
                t = (w * p)
                out.append(t.transpose(0, -1).contiguous().transpose(0, -1) if t.dim() > 1 else t)
        return out

    def plain_grads(cur):
        return [g.contiguous() for g in odd_grads(cur)]

    assert any(not g.is_contiguous() for g in odd_grads([b.requires_grad_(True) for b in base]))
    got = _run(base, odd_grads, douts, recipe, transform)
    want = _run(base, plain_grads, douts, recipe, transform)
    for a, b in zip(got[0] + got[1], want[0] + want[1]):
        assert a.shape == b.shape and torch.equal(a, b)


@pytest.mark.parametrize("transform", [False, True], ids=["whole_step", "scale_by_adam"])
def test_channels_last_parameters(transform):
    """channels_last parameters (moments inherit the strides) with row-major gradients and row-major incoming
    meta-gradients: logical results equal the all-row-major run."""
    base = _leaves()
    cl = [p.contiguous(memory_format=torch.channels_last) if p.dim() == 4 else p for p in base]
    assert not cl[0].is_contiguous() and cl[0].is_contiguous(memory_format=torch.channels_last)
    ws = [torch.rand_like(p, memory_format=torch.contiguous_format) + 0.5 for p in base]
    douts = [torch.randn_like(p, memory_format=torch.contiguous_format) for p in base]

    def grads(cur):
        return [(w * p).contiguous() for w, p in zip(ws, cur)]   # row-major, whatever the parameter's strides

    got = _run(cl, grads, douts, dict(), transform)
    want = _run(base, grads, douts, dict(), transform)
    for a, b in zip(got[0] + got[1], want[0] + want[1]):
        assert a.shape == b.shape and torch.equal(a, b)


def test_inplace_step_rejects_mismatched_state_and_relays_gradient():
    import torchopt_b200 as tb
    p = torch.randn(6, 5, device="cuda")
    mu, nu = torch.zeros(5, 6, device="cuda").t(), torch.zeros(6, 5, device="cuda")
    g = torch.randn(6, 5, device="cuda")
    with pytest.raises(RuntimeError, match="in-place step"):
        tb.adam_op.fused_step([p], [g], [mu], [nu], 0.9, 0.999, 1e-8, 0.0, -1e-3, [1], True)
    # a column-major gradient for a row-major parameter: re-laid-out for the kernel, and the reference's side effect
    # (the gradient tensor becomes the scaled update) lands in the caller's tensor in ITS layout
    g_t = g.t().contiguous().t()
    p1, p2 = p.clone(), p.clone()
    mu1, nu1, mu2, nu2 = (torch.zeros_like(p) for _ in range(4))
    ga, gb = g.clone(), g_t.clone(memory_format=torch.preserve_format)
    assert not gb.is_contiguous()
    tb.adam_op.fused_step([p1], [ga], [mu1], [nu1], 0.9, 0.999, 1e-8, 0.0, -1e-3, [1], True)
    tb.adam_op.fused_step([p2], [gb], [mu2], [nu2], 0.9, 0.999, 1e-8, 0.0, -1e-3, [1], True)
    torch.cuda.synchronize()
    assert torch.equal(p1, p2) and torch.equal(mu1, mu2) and torch.equal(nu1, nu2) and torch.equal(ga, gb)


def test_flat_alloc_does_not_pin_large_outputs():
    """62-leaf fused_forward (ResNet-18 leaf sizes): small outputs share one pooled buffer, large ones own their
    storage -- dropping everything but the smallest output must release (almost) all of the call's memory."""
    import torchopt_b200 as tb
    import bench
    sizes = bench.RESNET18_LEAVES
    g = [torch.randn(k, device="cuda") for k in sizes]
    mu = [torch.zeros(k, device="cuda") for k in sizes]
    nu = [torch.zeros(k, device="cuda") for k in sizes]
    torch.cuda.synchronize()
    before = torch.cuda.memory_allocated()
    new_mu, new_nu, new_u = tb.adam_op.fused_forward(g, mu, nu, 0.9, 0.999, 1e-8, 0.0, [1] * len(sizes), False)
    torch.cuda.synchronize()
    full = torch.cuda.memory_allocated() - before
    assert full >= 3 * 4 * sum(sizes)
    k = min(range(len(sizes)), key=lambda i: sizes[i])
    keep = new_u[k]
    del new_mu, new_nu, new_u
    left = torch.cuda.memory_allocated() - before
    pooled = sum(3 * (4 * s + 128) for s in sizes if 4 * s < (256 << 10))
    assert left <= pooled + (1 << 20), (left, pooled, full)
    assert left < full // 20
    assert keep.numel() == sizes[k]
