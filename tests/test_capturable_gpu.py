"""Capturable in-place Adam (`Adam(..., capturable=True)`, `FlatAdam(..., capturable=True)`,
`b200_adam_step_inplace_capturable`): step counts on the device, bias corrections from host-computed tables.
Must be bit-identical to the host-scalar path at every step count -- inside the tables, at their last entry and far
beyond it -- and a training iteration recorded in a CUDA graph must advance the optimizer by one real step per replay."""
from __future__ import annotations

import numpy as np
import pytest

from conftest import assert_bits_equal

torch = pytest.importorskip("torch")
from torch import nn  # noqa: E402

pytestmark = pytest.mark.gpu
SIZES = [5, 64, 321, 4097, 36_864]


def test_bias_tables_match_the_reference_expression():
    """The tables hold float(1 / (1 - pow(b, count))) -- the reference launchers' host expression
    (adam_op_impl_cuda.cu:73-75) -- for count = 1..len, and end exactly where both corrections have become 1."""
    import math

    import torchopt_b200 as tb
    for family, dt in ((0, np.float32), (1, np.float64)):
        for b1, b2 in ((0.9, 0.999), (0.5, 0.6), (0.0, 0.3)):
            c1, c2 = (t.numpy() for t in tb.adam_op.bias_tables(family, b1, b2))
            n = len(c1)
            want1 = np.array([dt(1 / (1 - math.pow(b1, c))) for c in range(1, n + 1)], dtype=dt)
            want2 = np.array([dt(1 / (1 - math.pow(b2, c))) for c in range(1, n + 1)], dtype=dt)
            assert_bits_equal(c1, want1, "c1 table")
            assert_bits_equal(c2, want2, "c2 table")
            assert c1[-1] == 1 and c2[-1] == 1 and (n == 1 or c1[-2] != 1 or c2[-2] != 1)
            for c in (n + 1, 10 * n, 10**9):
                assert dt(1 / (1 - math.pow(b1, c))) == 1 and dt(1 / (1 - math.pow(b2, c))) == 1
    with pytest.raises(RuntimeError, match="2\\^24"):
        tb.adam_op.bias_tables(0, 0.9, 0.9999999)


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_capturable_abi_vs_oracle_at_any_count(c_oracle, tag):
    """Raw C ABI: device count + tables vs the oracle's whole-step composition, for counts inside the table, at its end
    and far beyond it; gradients overwritten or left alone."""
    from oracle import adam_oracle as ao
    from torchopt_b200 import abi
    import torchopt_b200 as tb
    dt, npdt = (abi.F32, np.float32) if tag == "f32" else (abi.F64, np.float64)
    b1, b2, eps, er, lr = 0.9, 0.999, 1e-8, 1e-8, 3e-3
    c1, c2 = (t.cuda() for t in tb.adam_op.bias_tables(1 if tag == "f64" else 0, b1, b2))
    T = c1.numel()
    rng = np.random.default_rng(5)
    s = torch.cuda.current_stream().cuda_stream
    lib = abi.load()
    for n in (1, 1023, 70_001):
        p = (rng.standard_normal(n) * 0.3).astype(npdt)
        g = (rng.standard_normal(n) * 1e-2).astype(npdt)
        mu = (rng.standard_normal(n) * 1e-2).astype(npdt)
        nu = (rng.random(n) * 1e-4).astype(npdt)
        for count in (1, 7, T - 1, T, T + 1, 10**6, 2**40):
            w_p, w_mu, w_nu, w_us = ao.fused_step_forward(c_oracle, p, g, mu, nu, b1, b2, eps, er, count, -lr)
            for overwrite in (0, 1):
                P, G, MU, NU = (torch.from_numpy(a.copy()).cuda() for a in (p, g, mu, nu))
                cnt = torch.tensor(count, dtype=torch.int64, device="cuda")
                rc = lib.b200_adam_step_inplace_capturable(
                    dt, 1, abi._ptrs([P.data_ptr()]), abi._ptrs([G.data_ptr()]), abi._ptrs([MU.data_ptr()]),
                    abi._ptrs([NU.data_ptr()]), abi._arr(abi.c_int64, [n]), abi._ptrs([cnt.data_ptr()]), c1.data_ptr(),
                    c2.data_ptr(), T, b1, b2, eps, er, -lr, 0.0, 0.0, 0, overwrite, s)
                assert rc == 0
                torch.cuda.synchronize()
                what = f"capturable n={n} count={count} overwrite={overwrite}"
                assert_bits_equal(P.cpu().numpy(), w_p, what + " p")
                assert_bits_equal(MU.cpu().numpy(), w_mu, what + " mu")
                assert_bits_equal(NU.cpu().numpy(), w_nu, what + " nu")
                assert_bits_equal(G.cpu().numpy(), w_us if overwrite else g, what + " g")


class Elementwise(nn.Module):
    def __init__(self, sizes, dtype=torch.float32):
        super().__init__()
        self.ps = nn.ParameterList([nn.Parameter((torch.randn(k, device="cuda") * 0.3).to(dtype)) for k in sizes])

    def forward(self, ws):
        return torch.stack([(w * p * p).sum() for w, p in zip(ws, self.ps)]).sum()


@pytest.mark.parametrize("cls_name,dtype,betas,steps", [
    ("FlatAdam", torch.float32, (0.5, 0.6), 45),        # table of 33 entries: crossed during the run
    ("FlatAdamW", torch.float32, (0.9, 0.999), 12),
    ("Adam", torch.float32, (0.5, 0.6), 45),
    ("AdamW", torch.float64, (0.9, 0.999), 8),
    ("FlatAdam", torch.bfloat16, (0.9, 0.999), 8),
])
def test_capturable_equals_host_scalar_path_eagerly(cls_name, dtype, betas, steps):
    import torchopt_b200 as tb
    torch.manual_seed(11)
    net_a, net_b = Elementwise(SIZES, dtype), Elementwise(SIZES, dtype)
    net_b.load_state_dict(net_a.state_dict())
    cls = getattr(tb, cls_name)
    kw = dict(lr=1e-2, betas=betas, eps_root=1e-8, weight_decay=0.01)
    opt_a, opt_b = cls(net_a.parameters(), **kw), cls(net_b.parameters(), capturable=True, **kw)
    ws = [(torch.rand(k, device="cuda") + 0.5).to(dtype) for k in SIZES]
    for k in range(steps):
        for net, opt in ((net_a, opt_a), (net_b, opt_b)):
            opt.zero_grad()
            net(ws).backward()
            if cls_name in ("Adam", "AdamW") and k % 5 == 1:
                list(net.parameters())[2].grad = None          # a leaf that skips this step: its count stays behind
            opt.step()
        for p, q in zip(net_a.parameters(), net_b.parameters()):
            assert torch.equal(p, q), f"step {k}"
    st = next(s for s in opt_b.state if isinstance(s, tb.ScaleByAdamState))
    counts = [int(c) for c in st.count]
    if cls_name in ("Adam", "AdamW"):
        skipped = len([k for k in range(steps) if k % 5 == 1])
        assert counts == [steps, steps, steps - skipped, steps, steps]
    else:
        assert counts == [steps]


@pytest.mark.parametrize("flat", [True, False])
def test_training_iteration_recorded_in_a_cuda_graph(flat):
    """zero_grad + forward + backward + optimizer step in ONE graph: N replays == N eager iterations, bit for bit."""
    import torchopt_b200 as tb
    torch.manual_seed(12)
    net_a, net_b = Elementwise(SIZES), Elementwise(SIZES)
    net_b.load_state_dict(net_a.state_dict())
    cls = tb.FlatAdam if flat else tb.Adam
    kw = dict(lr=1e-2, betas=(0.7, 0.8), eps_root=1e-8)
    opt_a, opt_b = cls(net_a.parameters(), **kw), cls(net_b.parameters(), capturable=True, **kw)
    w = [torch.rand(k, device="cuda") + 0.5 for k in SIZES]

    def iteration(net, opt, ws):
        opt.zero_grad()
        loss = net(ws)
        loss.backward()
        opt.step()
        return loss.detach()

    warmup, replays = 2, 9
    step = tb.graphs.capture(lambda *ws: iteration(net_b, opt_b, ws), *w, warmup=warmup)
    losses_b = [step().clone() for _ in range(replays)]
    losses_a = [iteration(net_a, opt_a, w) for _ in range(warmup + replays)]
    torch.cuda.synchronize()
    for p, q in zip(net_a.parameters(), net_b.parameters()):
        assert torch.equal(p, q)
    for a, b in zip(losses_a[warmup:], losses_b):
        assert torch.equal(a, b)
    st = next(s for s in opt_b.state if isinstance(s, tb.ScaleByAdamState))
    assert all(int(c) == warmup + replays for c in st.count)
    # new data through the static input slots
    w2 = [torch.rand(k, device="cuda") + 0.25 for k in SIZES]
    lb = step(*w2).clone()
    la = iteration(net_a, opt_a, w2)
    assert torch.equal(la, lb)
    for p, q in zip(net_a.parameters(), net_b.parameters()):
        assert torch.equal(p, q)


# ---------------------------------------------------------------------------------------------------------------
# capturable DIFFERENTIABLE step (round 2): MetaAdam / FuncAdam(capturable=True) over a persistent state
# ---------------------------------------------------------------------------------------------------------------
def test_bias_tables_ex_match_the_reference_expressions():
    """[c1, c2, o1, c2e]: the four host expressions of the reference launchers (adam_op_impl_cuda.cu:73-75,452-454), one
    rounding each; the tables end where all four have reached their limit values."""
    import math

    import torchopt_b200 as tb
    for family, dt in ((0, np.float32), (1, np.float64)):
        for b1, b2, er in ((0.9, 0.999, 0.0), (0.9, 0.999, 1e-8), (0.5, 0.6, 1e-3)):
            c1, c2, o1, c2e = (t.numpy() for t in tb.adam_op.bias_tables_ex(family, b1, b2, er))
            n = len(c1)
            cs = range(1, n + 1)
            assert_bits_equal(c1, np.array([dt(1 / (1 - math.pow(b1, c))) for c in cs], dtype=dt), "c1")
            assert_bits_equal(c2, np.array([dt(1 / (1 - math.pow(b2, c))) for c in cs], dtype=dt), "c2")
            assert_bits_equal(o1, np.array([dt(1 - math.pow(b1, c)) for c in cs], dtype=dt), "o1")
            assert_bits_equal(c2e, np.array([dt(1 / (1 - math.pow(b2, c) + er)) for c in cs], dtype=dt), "c2e")
            for c in (n + 1, 10 * n, 10**9):    # every larger count reproduces the last entry
                assert dt(1 / (1 - math.pow(b1, c))) == c1[-1] and dt(1 / (1 - math.pow(b2, c))) == c2[-1]
                assert dt(1 - math.pow(b1, c)) == o1[-1] and dt(1 / (1 - math.pow(b2, c) + er)) == c2e[-1]


@pytest.mark.parametrize("kw", [dict(), dict(weight_decay=1e-2), dict(weight_decay=1e-2, maximize=True)],
                         ids=["adam", "l2", "l2_max"])
@pytest.mark.parametrize("mrg", [False, True])
def test_capturable_differentiable_step_eager_equals_host_scalar_path(kw, mrg):
    """Eager (no graph): the device-count + table form gives the same parameters, moments and gradients as the default
    host-scalar form, step after step, including a leaf that is skipped in some steps (its count must not advance)."""
    import torchopt_b200 as tb
    gen = torch.Generator("cuda").manual_seed(21)
    p0 = [torch.randn(k, device="cuda", generator=gen) * 0.3 for k in SIZES]
    ws = [[torch.rand(k, device="cuda", generator=gen) + 0.5 for k in SIZES] for _ in range(5)]

    def run(capturable):
        opt = tb.FuncAdam(1e-2, eps_root=1e-8, moment_requires_grad=mrg, capturable=capturable, **kw)
        cur = [p.clone() for p in p0]
        out = []
        for k in range(5):
            leaves = [p.detach().requires_grad_(True) for p in cur]
            hyper = [w.clone().requires_grad_(True) for w in ws[k]]
            grads = [w * p for w, p in zip(hyper, leaves)]
            if k in (1, 3):
                grads[2] = None                       # skipped leaf: parameter, moments and count stay
            new_p = opt.apply_gradients(grads, leaves)
            outer = sum((q * q).sum() for q in new_p)
            meta = torch.autograd.grad(outer, hyper + leaves, allow_unused=True)
            cur = [q.detach() for q in new_p]
            tb.stop_gradient(opt)                     # the capturable state is carried by value: same cut here
            out.append((cur, meta))
        st = next(s for s in opt.state if isinstance(s, tb.ScaleByAdamState))
        torch.cuda.synchronize()
        return out, [m.detach().clone() for m in st.mu], [int(c) for c in st.count]

    a, mu_a, cnt_a = run(True)
    b, mu_b, cnt_b = run(False)
    assert cnt_a == cnt_b == [5, 5, 3, 5, 5]
    for (cur_a, meta_a), (cur_b, meta_b) in zip(a, b):
        for x, y in zip(cur_a + list(meta_a), cur_b + list(meta_b)):
            assert (x is None) == (y is None)
            if x is not None:
                assert torch.equal(x, y)
    for x, y in zip(mu_a, mu_b):
        assert torch.equal(x, y)


def test_differentiable_step_with_persistent_state_recorded_in_a_cuda_graph():
    """K replays of ONE captured differentiable step (persistent FuncAdam(capturable=True) state, parameters in static
    buffers, meta-gradient w.r.t. a hyper-parameter taken inside the graph) == K eager steps, bit for bit.  The
    reference cannot be recorded at all (count -> host on every call, accelerated_op/adam_op.py:98-101); without
    capturable=True this package refuses (test_flat_state_gpu.py) because the bias corrections would be frozen."""
    import torchopt_b200 as tb
    gen = torch.Generator("cuda").manual_seed(31)
    p0 = [torch.randn(k, device="cuda", generator=gen) * 0.3 for k in SIZES]
    w0 = [torch.rand(k, device="cuda", generator=gen) + 0.5 for k in SIZES]
    K, WARM = 6, 3     # graphs.capture runs the function WARM times eagerly before recording: those are real steps too

    def make():
        params = [p.clone() for p in p0]                                    # static parameter buffers
        hyper = [w.clone().requires_grad_(True) for w in w0]
        return params, hyper

    # eager reference: K + WARM + 1 (the recording itself does not execute) steps of the host-scalar optimizer
    params, hyper = make()
    opt = tb.FuncAdam(1e-2, eps_root=1e-8, weight_decay=1e-2)
    want = []
    for _ in range(WARM + K):
        leaves = [p.detach().requires_grad_(True) for p in params]
        new_p = opt.apply_gradients([w * p for w, p in zip(hyper, leaves)], leaves)
        meta = torch.autograd.grad(sum((q * q).sum() for q in new_p), hyper)
        with torch.no_grad():
            for p, q in zip(params, new_p):
                p.copy_(q)
        tb.stop_gradient(opt)
        want.append(([p.clone() for p in params], [m.clone() for m in meta]))

    params, hyper = make()
    opt_c = tb.FuncAdam(1e-2, eps_root=1e-8, weight_decay=1e-2, capturable=True)

    def one_step():
        leaves = [p.detach().requires_grad_(True) for p in params]
        new_p = opt_c.apply_gradients([w * p for w, p in zip(hyper, leaves)], leaves)
        meta = torch.autograd.grad(sum((q * q).sum() for q in new_p), hyper)
        with torch.no_grad():
            for p, q in zip(params, new_p):
                p.copy_(q)                                                  # persistent parameters: static addresses
        return meta

    step = tb.graphs.capture(one_step, warmup=WARM)
    got = []
    for _ in range(K):
        meta = step()
        torch.cuda.synchronize()
        got.append(([p.clone() for p in params], [m.clone() for m in meta]))
    for k in range(K):
        for x, y in zip(got[k][0] + got[k][1], want[WARM + k][0] + want[WARM + k][1]):
            assert torch.equal(x, y), f"replay {k} differs from eager step {WARM + k + 1}"
    st = next(s for s in opt_c.state if isinstance(s, tb.ScaleByAdamState))
    assert [int(c) for c in st.count] == [WARM + K] * len(SIZES)


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_capturable_differentiable_abi_vs_oracle_at_any_count(c_oracle, tag):
    """Raw C ABI of the capturable DIFFERENTIABLE step (b200_adam_step_forward_capturable / _backward_capturable): device
    count + four tables vs the oracle's whole-step forward and backward, for counts inside the tables, at their end and
    far beyond; full and partial presence of the incoming gradients; two leaves with different counts in one launch."""
    from oracle import adam_oracle as ao
    from torchopt_b200 import abi
    import torchopt_b200 as tb
    dt, npdt = (abi.F32, np.float32) if tag == "f32" else (abi.F64, np.float64)
    b1, b2, eps, er, lr = 0.9, 0.999, 1e-8, 1e-8, 3e-3
    tabs = [t.cuda() for t in tb.adam_op.bias_tables_ex(1 if tag == "f64" else 0, b1, b2, er)]
    T = tabs[0].numel()
    rng = np.random.default_rng(9)
    s = torch.cuda.current_stream().cuda_stream
    lib = abi.load()
    ptrs, arr = abi._ptrs, abi._arr
    sizes = (1023, 70_001)
    for counts in ((1, 7), (T - 1, T), (T + 1, 10**6), (2**40, 3)):
        leaves = []
        for n, count in zip(sizes, counts):
            a = dict(p=(rng.standard_normal(n) * 0.3).astype(npdt), g=(rng.standard_normal(n) * 1e-2).astype(npdt),
                     mu=(rng.standard_normal(n) * 1e-2).astype(npdt), nu=(rng.random(n) * 1e-4).astype(npdt),
                     dp=rng.standard_normal(n).astype(npdt), dme=(rng.standard_normal(n) * 1e-3).astype(npdt),
                     dne=(rng.standard_normal(n) * 1e-3).astype(npdt))
            a["g"][::11] = 0
            a["mu"][::11] = 0
            leaves.append((n, count, a, {k: torch.from_numpy(v).cuda() for k, v in a.items()}))
        cnt = torch.tensor(list(counts), dtype=torch.int64, device="cuda")
        cptr = [cnt.data_ptr(), cnt.data_ptr() + 8]
        outs = [{k: torch.full((n,), 7.0, device="cuda", dtype=torch.float64 if tag == "f64" else torch.float32)
                 for k in ("p2", "mu2", "nu2", "dg", "dmu", "dnu")} for n, _, _, _ in leaves]
        col = lambda k: ptrs([d[k].data_ptr() for _, _, _, d in leaves])        # noqa: E731
        ocol = lambda k: ptrs([o[k].data_ptr() for o in outs])                   # noqa: E731
        ns = arr(abi.c_int64, list(sizes))
        rc = lib.b200_adam_step_forward_capturable(dt, 2, col("p"), col("g"), col("mu"), col("nu"), ocol("p2"), ocol("mu2"),
                                                   ocol("nu2"), None, ns, ptrs(cptr), tabs[0].data_ptr(), tabs[1].data_ptr(),
                                                   T, b1, b2, eps, er, -lr, 0.0, 0.0, 0, s)
        assert rc == 0
        torch.cuda.synchronize()
        want_fwd = []
        for (n, count, a, _), o in zip(leaves, outs):
            w_p, w_mu, w_nu, _ = ao.fused_step_forward(c_oracle, a["p"], a["g"], a["mu"], a["nu"], b1, b2, eps, er, count, -lr)
            want_fwd.append((w_mu, w_nu))
            what = f"capturable step forward n={n} count={count}"
            assert_bits_equal(o["p2"].cpu().numpy(), w_p, what + " p'")
            assert_bits_equal(o["mu2"].cpu().numpy(), w_mu, what + " mu'")
            assert_bits_equal(o["nu2"].cpu().numpy(), w_nu, what + " nu'")
        for hme, hne in ((1, 1), (0, 0), (1, 0)):
            dme = ptrs([d["dme"].data_ptr() if hme else None for _, _, _, d in leaves])
            dne = ptrs([d["dne"].data_ptr() if hne else None for _, _, _, d in leaves])
            rc = lib.b200_adam_step_backward_capturable(
                dt, 2, col("dp"), None, ocol("mu2"), ocol("nu2"), col("g"), dme, dne, ocol("dg"), ocol("dmu"), ocol("dnu"),
                None, None, ns, ptrs(cptr), tabs[0].data_ptr(), tabs[1].data_ptr(), tabs[2].data_ptr(), tabs[3].data_ptr(), T,
                b1, b2, eps, er, -lr, 0.0, 0.0, 0, s)
            assert rc == 0
            torch.cuda.synchronize()
            for (n, count, a, _), o, (w_mu, w_nu) in zip(leaves, outs, want_fwd):
                want = ao.fused_step_backward(c_oracle, a["dp"], w_mu, w_nu, a["g"], a["dme"] if hme else None,
                                              a["dne"] if hne else None, b1, b2, eps, er, count, -lr)
                for k, name in enumerate(("dg", "dmu", "dnu")):
                    assert_bits_equal(o[name].cpu().numpy(), want[k],
                                      f"capturable step backward {name} n={n} count={count} pattern={(hme, hne)}")
    # argument errors: a missing table / count pointer is refused, not dereferenced
    assert lib.b200_adam_step_forward_capturable(dt, 2, col("p"), col("g"), col("mu"), col("nu"), ocol("p2"), ocol("mu2"),
                                                 ocol("nu2"), None, ns, ptrs([cptr[0], None]), tabs[0].data_ptr(),
                                                 tabs[1].data_ptr(), T, b1, b2, eps, er, -lr, 0.0, 0.0, 0, s) == -2
    assert lib.b200_adam_step_backward_capturable(dt, 2, col("dp"), None, ocol("mu2"), ocol("nu2"), col("g"), dme, dne,
                                                  ocol("dg"), ocol("dmu"), ocol("dnu"), None, None, ns, ptrs(cptr),
                                                  tabs[0].data_ptr(), tabs[1].data_ptr(), None, tabs[3].data_ptr(), T, b1, b2,
                                                  eps, er, -lr, 0.0, 0.0, 0, s) == -2


def test_online_meta_gradient_example_runs_and_counts_on_the_device():
    """examples/online_meta_gradient.py (also `extra.online_meta_gradient` of bench.py): every eager call and every graph
    replay is one real optimizer step -- the device step count says so."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("b200_example_online", os.path.join(root, "examples", "online_meta_gradient.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rec = mod.measure(log2n=14, steps=7)
    assert rec["device_step_count_after"] == 3 + 3 + 7      # capture warm-ups + timing warm-ups + timed replays
    assert rec["cuda_graph_capturable_ms"] > 0 and rec["eager_default_ms"] > 0
