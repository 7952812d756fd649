"""extract_state_dict / recover_state_dict / stop_gradient (torchopt_b200.utils; reference torchopt/utils.py:56-349,
used by its MAML loops, examples/few-shot/maml_omniglot.py:136-165).  Module handling is pure reference shuffling and is
tested on CPU; the optimizer round trips need the kernels and run on the GPU."""
from __future__ import annotations

import pytest
import torch
from torch import nn

gpu = pytest.mark.gpu


def _net():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(4, 3), nn.BatchNorm1d(3), nn.Tanh(), nn.Linear(3, 2))


def test_module_state_roundtrip_by_reference():
    import torchopt_b200 as tb
    net = _net()
    originals = list(net.parameters())
    state = tb.extract_state_dict(net)
    assert isinstance(state, tb.ModuleState) and len(state.params) == 3 and len(state.buffers) == 1
    assert all(a is b for a, b in zip((v for d in state.params for v in d.values()), originals))
    # what a differentiable optimizer does: re-bind the slots to non-leaf tensors
    for m in net.modules():
        for k, p in list(m._parameters.items()):
            m._parameters[k] = p * 2.0
    assert all(p.grad_fn is not None for p in net.parameters())
    tb.recover_state_dict(net, state)
    assert all(a is b for a, b in zip(net.parameters(), originals))
    with pytest.raises(TypeError):
        tb.recover_state_dict(net, ("not", "a", "state"))
    with pytest.raises(TypeError):
        tb.extract_state_dict(object())


def test_module_state_copy_modes_and_buffers():
    import torchopt_b200 as tb
    net = _net()
    by_copy = tb.extract_state_dict(net, by="copy")
    w = by_copy.params[0]["weight"]
    assert w is not net[0].weight and w.grad_fn is not None and torch.equal(w, net[0].weight)
    deep = tb.extract_state_dict(net, by="deepcopy")
    w = deep.params[0]["weight"]
    assert isinstance(w, nn.Parameter) and w.grad_fn is None and w.requires_grad and w is not net[0].weight
    ref = tb.extract_state_dict(net, detach_buffers=True)
    rm = ref.buffers[0]["running_mean"]
    assert rm is not net[1].running_mean and torch.equal(rm, net[1].running_mean)
    net(torch.randn(8, 4))                       # training-mode BatchNorm updates its buffers in place
    assert not torch.equal(rm, net[1].running_mean)
    tb.recover_state_dict(net, ref)              # detach_buffers: the module gets fresh clones of the saved buffers
    assert torch.equal(net[1].running_mean, rm) and net[1].running_mean is not rm
    none = tb.extract_state_dict(net, with_buffers=False)
    assert none.buffers == ()
    with pytest.raises(ValueError):
        tb.extract_state_dict(net, by="sideways")


def test_stop_gradient_cuts_in_place_and_keeps_requires_grad():
    import torchopt_b200 as tb
    a = torch.randn(3, requires_grad=True)
    tree = {"x": [a * 2, (a + 1,)], "leaf": a, "const": torch.ones(2)}
    tb.stop_gradient(tree)
    assert tree["x"][0].grad_fn is None and tree["x"][0].requires_grad
    assert tree["x"][1][0].grad_fn is None and tree["leaf"] is a and not tree["const"].requires_grad
    net = _net()
    net[0]._parameters["weight"] = net[0].weight * 1.0
    tb.stop_gradient(net)
    assert net[0]._parameters["weight"].grad_fn is None and net[0]._parameters["weight"].requires_grad


@gpu
@pytest.mark.parametrize("flat", [False, True])
def test_maml_loop_with_extract_recover(flat):
    """The reference's MAML loop shape: extract both states once, recover them after every task.  Every task must see
    the same meta-parameters and a fresh Adam state, so two identical tasks give identical inner results."""
    import torch.nn.functional as F

    import torchopt_b200 as tb
    dev = torch.device("cuda")
    torch.manual_seed(1)
    torch.backends.cuda.matmul.allow_tf32 = False
    net = nn.Sequential(nn.Linear(6, 16), nn.Tanh(), nn.Linear(16, 3)).to(dev)
    meta = list(net.parameters())
    x, y = torch.randn(10, 6, device=dev), torch.randint(0, 3, (10,), device=dev)
    inner = (tb.FlatMetaAdam if flat else tb.MetaAdam)(net, lr=1e-2)
    net_state = tb.extract_state_dict(net)
    opt_state = tb.extract_state_dict(inner)
    results = []
    for _ in range(2):
        for _ in range(3):
            inner.step(F.cross_entropy(net(x), y))
        loss = F.cross_entropy(net(x), y)
        grads = torch.autograd.grad(loss, meta)
        results.append((loss.detach().clone(), [g.clone() for g in grads]))
        tb.recover_state_dict(net, net_state)
        tb.recover_state_dict(inner, opt_state)
    assert torch.equal(results[0][0], results[1][0])
    for a, b in zip(results[0][1], results[1][1]):
        assert torch.equal(a, b)
    if not flat:
        assert all(a is b for a, b in zip(net.parameters(), meta))


@gpu
@pytest.mark.parametrize("flat", [False, True])
def test_stop_gradient_truncates_the_unroll(flat):
    """After stop_gradient(net) + stop_gradient(optimizer) nothing flows to the meta-parameters any more, while the
    unroll continues from the same values (truncated back-propagation, reference tests/test_utils.py:26-70)."""
    import torch.nn.functional as F

    import torchopt_b200 as tb
    dev = torch.device("cuda")
    torch.manual_seed(2)
    net = nn.Sequential(nn.Linear(6, 8), nn.Tanh(), nn.Linear(8, 3)).to(dev)
    meta = list(net.parameters())
    x, y = torch.randn(10, 6, device=dev), torch.randint(0, 3, (10,), device=dev)
    inner = (tb.FlatMetaAdam if flat else tb.MetaAdam)(net, lr=1e-2)
    inner.step(F.cross_entropy(net(x), y))
    before = [p.detach().clone() for p in net.parameters()]
    if flat:
        tb.stop_gradient(inner)            # cuts the flat buffer and the moments, re-binds the module's views
    else:
        tb.stop_gradient(net)
        tb.stop_gradient(inner)
    for a, b in zip(net.parameters(), before):
        assert torch.equal(a.detach(), b)
    inner.step(F.cross_entropy(net(x), y))
    grads = torch.autograd.grad(F.cross_entropy(net(x), y), meta, allow_unused=True)
    assert all(g is None for g in grads), "the graph no longer reaches the meta-parameters"
    cut_points = [inner.flat_p] if flat else None
    if not flat:
        # the truncated unroll is still differentiable w.r.t. the cut tensors
        inner2 = tb.MetaAdam(net, lr=1e-2)
        leaves = list(net.parameters())
        inner2.step(F.cross_entropy(net(x), y))
        g2 = torch.autograd.grad(F.cross_entropy(net(x), y), leaves, allow_unused=True)
        assert all(g is not None for g in g2)
    assert cut_points is None or cut_points[0].requires_grad
