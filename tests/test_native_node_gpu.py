"""The native autograd nodes (torch_shim.cpp: FusedNode, StepNode, and FlatStepNode -- the default path of the optimizer
classes, flat moment buffers) against the Python autograd.Functions they replace (AdamOp.FusedOp, AdamStep.Op, per-leaf
state): same values, same gradients, same None patterns, bit for bit -- they call the same kernels with the same
arguments; only the bookkeeping moves from the interpreter to C++ and from 4L tensors to 2L+2."""
from __future__ import annotations

import contextlib

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

SIZES = [5, 64, 321, 4097, 36_864]


@contextlib.contextmanager
def python_nodes():
    from torchopt_b200 import optim
    from torchopt_b200.accelerated_op import adam_op, adam_step
    adam_op.NATIVE_NODE = adam_step.NATIVE_NODE = False
    optim.FLAT_STATE = False    # per-leaf state: the optimizers go through AdamStep.Op instead of the flat-state node
    try:
        yield
    finally:
        adam_op.NATIVE_NODE = adam_step.NATIVE_NODE = True
        optim.FLAT_STATE = True


def same(a, b):
    if a is None or b is None:
        return a is None and b is None
    return torch.equal(a, b)


@pytest.mark.parametrize("mrg", [False, True])
@pytest.mark.parametrize("variant", ["adam", "adam_l2", "adamw_max"])
def test_step_node_unroll_bitwise(mrg, variant):
    import torchopt_b200 as tb
    dev = torch.device("cuda")
    torch.manual_seed(3)
    params = [(torch.randn(k, device=dev) * 0.3).requires_grad_(True) for k in SIZES]
    ws = [[torch.rand(k, device=dev) + 0.5 for k in SIZES] for _ in range(4)]
    cls, kw = {"adam": (tb.FuncAdam, {}), "adam_l2": (tb.FuncAdam, {"weight_decay": 0.05}),
               "adamw_max": (tb.FuncAdamW, {"weight_decay": 0.05, "maximize": True})}[variant]

    def run():
        opt = cls(1e-2, eps_root=1e-8, moment_requires_grad=mrg, **kw)
        cur = list(params)
        l0 = tb.abi.kernel_launches()
        for k in range(3):
            grads = [w * p for w, p in zip(ws[k], cur)]
            grads[1] = None                                    # a leaf without gradient passes through
            cur = opt.apply_gradients(grads, cur)
        st = next(s for s in opt.state if isinstance(s, tb.ScaleByAdamState))
        outer = torch.stack([(w * p * p).sum() for w, p in zip(ws[3], cur)]).sum() + sum((m * m).sum() for m in st.mu)
        meta = torch.autograd.grad(outer, params, allow_unused=True)
        return [p.detach() for p in cur], meta, tb.abi.kernel_launches() - l0

    cur_n, meta_n, launches_n = run()
    with python_nodes():
        cur_p, meta_p, launches_p = run()
    assert launches_n == launches_p == 6
    assert all(same(a, b) for a, b in zip(cur_n, cur_p))
    assert all(same(a, b) for a, b in zip(meta_n, meta_p))
    assert cur_n[1] is not None and torch.equal(cur_n[1], params[1].detach())


@pytest.mark.parametrize("mrg", [False, True])
def test_fused_node_chained_path_bitwise(mrg):
    """scale_by_accelerated_adam (FusedNode) + scale + apply_updates through tb.adam, incl. a None gradient and a leaf
    whose update is never used by the outer loss (its incoming gradient is undefined in the backward)."""
    import torchopt_b200 as tb
    dev = torch.device("cuda")
    torch.manual_seed(4)
    params = [(torch.randn(k, device=dev) * 0.3).requires_grad_(True) for k in SIZES]
    ws = [[torch.rand(k, device=dev) + 0.5 for k in SIZES] for _ in range(3)]

    def run():
        opt = tb.adam(1e-2, eps_root=1e-8, moment_requires_grad=mrg)
        cur, state = list(params), opt.init(params)
        for k in range(3):
            grads = [w * p for w, p in zip(ws[k], cur)]
            grads[2] = None
            updates, state = opt.update(grads, state, params=cur, inplace=False)
            cur = tb.apply_updates(cur, updates, inplace=False)
        outer = torch.stack([(p * p).sum() for p in cur[:-1]]).sum()       # the last leaf never reaches the loss
        meta = torch.autograd.grad(outer, params, allow_unused=True)
        return [p.detach() for p in cur], meta

    cur_n, meta_n = run()
    with python_nodes():
        cur_p, meta_p = run()
    assert all(same(a, b) for a, b in zip(cur_n, cur_p))
    assert all(same(a, b) for a, b in zip(meta_n, meta_p))
    assert meta_n[-1] is None and meta_n[0] is not None


def test_single_leaf_adamop_call_and_no_grad_mode():
    """AdamOp.__call__ (one leaf, reference signature) through the native node; and under no_grad / with inputs that do
    not require grad no graph is built."""
    import torchopt_b200 as tb
    dev = torch.device("cuda")
    torch.manual_seed(5)
    g = (torch.randn(1000, device=dev) * 1e-2).requires_grad_(True)
    mu, nu = torch.zeros(1000, device=dev), torch.zeros(1000, device=dev)
    op = tb.AdamOp(eps_root=1e-8, inplace=False)
    new_mu, new_nu, u = op(mu, nu, g, 1)
    assert u.grad_fn is not None and new_mu.grad_fn is not None
    (dg,) = torch.autograd.grad((u * u).sum() + new_nu.sum(), [g])
    with python_nodes():
        new_mu_p, new_nu_p, u_p = op(mu, nu, g, 1)
        (dg_p,) = torch.autograd.grad((u_p * u_p).sum() + new_nu_p.sum(), [g])
    assert torch.equal(u, u_p) and torch.equal(new_mu, new_mu_p) and torch.equal(dg, dg_p)
    with torch.no_grad():
        _, _, u2 = op(mu, nu, g, 1)
    assert u2.grad_fn is None and not u2.requires_grad and torch.equal(u2, u)
    _, _, u3 = op(mu, nu, g.detach(), 1)
    assert u3.grad_fn is None
