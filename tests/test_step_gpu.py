"""GPU parity of the whole-step kernels (Adam + scale(-lr) + apply_updates in one launch):
through the C ABI vs the oracle composition, and through MetaAdam / Adam vs the chained path run with torch ops."""
from __future__ import annotations

import numpy as np
import pytest

from conftest import assert_bits_equal

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

B1, B2, EPS = 0.9, 0.999, 1e-8
NP = {"f32": np.float32, "f64": np.float64}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


def inputs(n, dt, seed):
    rng = np.random.default_rng(seed)
    p = (rng.standard_normal(n) * 0.3).astype(dt)
    g = (rng.standard_normal(n) * 1e-2).astype(dt)
    mu = (rng.standard_normal(n) * 1e-2).astype(dt)
    nu = (rng.random(n) * 1e-4).astype(dt)
    dp = rng.standard_normal(n).astype(dt)
    dus = rng.standard_normal(n).astype(dt)
    dme = (rng.standard_normal(n) * 1e-3).astype(dt)
    dne = (rng.standard_normal(n) * 1e-3).astype(dt)
    g[::13] = 0
    mu[::13] = 0
    return p, g, mu, nu, dp, dus, dme, dne


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_step_kernels_vs_oracle_composition(c_oracle, tag):
    from oracle import adam_oracle as ao
    from torchopt_b200 import abi
    dt = abi.F32 if tag == "f32" else abi.F64
    s = torch.cuda.current_stream().cuda_stream
    lr = 3e-3
    for n in (1, 7, 8, 9, 1023, 4096, 4097, 8193, 70_001):
        p, g, mu, nu, dp, dus, dme, dne = inputs(n, NP[tag], n)
        for er, cnt in ((0.0, 1), (1e-8, 4)):
            P, G, MU, NU, DP, DUS, DME, DNE = map(dev, (p, g, mu, nu, dp, dus, dme, dne))
            o = [torch.empty_like(P) for _ in range(4)]
            w_p, w_mu, w_nu, w_us = ao.fused_step_forward(c_oracle, p, g, mu, nu, B1, B2, EPS, er, cnt, -lr)
            for with_us in (True, False):
                for t in o:
                    t.fill_(7.0)
                abi.step_forward(dt, [P.data_ptr()], [G.data_ptr()], [MU.data_ptr()], [NU.data_ptr()],
                                 [o[0].data_ptr()], [o[1].data_ptr()], [o[2].data_ptr()],
                                 [o[3].data_ptr()] if with_us else None, [n], [cnt], B1, B2, EPS, er, -lr, s)
                assert_bits_equal(host(o[0]), w_p, f"step p' n={n}")
                assert_bits_equal(host(o[1]), w_mu, f"step mu' n={n}")
                assert_bits_equal(host(o[2]), w_nu, f"step nu' n={n}")
                if with_us:
                    assert_bits_equal(host(o[3]), w_us, f"step us n={n}")
                else:
                    assert (host(o[3]) == 7.0).all()
            # in place
            a, b, c, d = P.clone(), G.clone(), MU.clone(), NU.clone()
            abi.step_inplace(dt, [a.data_ptr()], [b.data_ptr()], [c.data_ptr()], [d.data_ptr()], [n], [cnt], B1, B2, EPS,
                             er, -lr, s)
            for got, want, name in ((a, w_p, "p"), (b, w_us, "g<-us"), (c, w_mu, "mu"), (d, w_nu, "nu")):
                assert_bits_equal(host(got), want, f"step_inplace {name} n={n}")
            # in place, gradients left untouched (b200_adam_step_inplace_ex, overwrite_grads = 0) and overwritten (= 1)
            for overwrite in (False, True):
                a, b, c, d = P.clone(), G.clone(), MU.clone(), NU.clone()
                abi.step_inplace(dt, [a.data_ptr()], [b.data_ptr()], [c.data_ptr()], [d.data_ptr()], [n], [cnt], B1, B2,
                                 EPS, er, -lr, s, overwrite_grads=overwrite)
                for got, want, name in ((a, w_p, "p"), (b, w_us if overwrite else g, "g"), (c, w_mu, "mu"), (d, w_nu, "nu")):
                    assert_bits_equal(host(got), want, f"step_inplace_ex(overwrite={overwrite}) {name} n={n}")
            # backward: FULL, LAST and GENERIC presence patterns (incl. a gradient on the scaled update)
            M2, V2 = dev(w_mu), dev(w_nu)
            for hdp, hdus, hme, hne in ((1, 0, 1, 1), (1, 0, 0, 0), (1, 1, 1, 0), (0, 1, 0, 1), (0, 0, 1, 1), (1, 0, 0, 1)):
                want = ao.fused_step_backward(c_oracle, dp if hdp else None, w_mu, w_nu, g, dme if hme else None,
                                              dne if hne else None, B1, B2, EPS, er, cnt, -lr, dus=dus if hdus else None)
                outs = [torch.full_like(P, 7.0) for _ in range(3)]
                abi.step_backward(dt, [DP.data_ptr() if hdp else None], [DUS.data_ptr() if hdus else None],
                                  [M2.data_ptr()], [V2.data_ptr()], [G.data_ptr()], [DME.data_ptr() if hme else None],
                                  [DNE.data_ptr() if hne else None], [outs[0].data_ptr()], [outs[1].data_ptr()],
                                  [outs[2].data_ptr()], [n], [cnt], B1, B2, EPS, er, -lr, s)
                for k in range(3):
                    got = host(outs[k])
                    if want[k] is None:
                        assert (got == 0).all()
                    else:
                        assert_bits_equal(got, want[k], f"step bwd out{k} n={n} pattern={(hdp, hdus, hme, hne)}")


def _mlp(dtype=torch.float32):
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(20, 33), torch.nn.Tanh(), torch.nn.Linear(33, 7)).to(dtype).cuda()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("mrg", [True, False])
def test_meta_adam_whole_step_equals_chained_path(dtype, mrg):
    """MetaAdam with the single whole-step launch vs the chained transformations + torch mul/add: parameters and
    meta-gradients bit-identical; Adam-kernel launches per inner step stay 1, torch's 2 kernels per leaf disappear."""
    import torchopt_b200 as tb
    from torchopt_b200 import optim
    x = torch.randn(16, 20, device="cuda", dtype=dtype, generator=torch.Generator("cuda").manual_seed(1))
    y = torch.randn(16, 7, device="cuda", dtype=dtype, generator=torch.Generator("cuda").manual_seed(2))
    out = {}
    for fuse in (True, False):
        optim.FUSE_STEP = fuse
        try:
            net = _mlp(dtype)
            init = list(net.parameters())
            opt = tb.MetaAdam(net, lr=1e-2, moment_requires_grad=mrg, eps_root=1e-8)
            before = tb.abi.kernel_launches()
            for _ in range(4):
                opt.step(((net(x) - y) ** 2).mean())
            launches = tb.abi.kernel_launches() - before
            outer = ((net(x) - y) ** 2).mean()
            meta = torch.autograd.grad(outer, init)
            torch.cuda.synchronize()
            out[fuse] = ([p.detach().cpu().numpy() for p in opt.params()], [m.cpu().numpy() for m in meta], launches,
                         [m.detach().cpu().numpy() for m in opt.state[0].mu], [int(c) for c in opt.state[0].count])
        finally:
            optim.FUSE_STEP = True
    assert out[True][2] == 4 and out[False][2] == 4
    assert out[True][4] == out[False][4] == [4, 4, 4, 4]
    for a, b in zip(out[True][0], out[False][0]):
        assert_bits_equal(a, b, "unrolled parameters")
    for a, b in zip(out[True][3], out[False][3]):
        assert_bits_equal(a, b, "first moments")
    for a, b in zip(out[True][1], out[False][1]):
        assert_bits_equal(a, b, "meta-gradients")


def test_inplace_adam_whole_step_equals_chained_path():
    import torchopt_b200 as tb
    from torchopt_b200 import optim
    res = {}
    for fuse in (True, False):
        optim.FUSE_STEP = fuse
        try:
            net = _mlp()
            opt = tb.Adam(net.parameters(), lr=1e-3, eps_root=1e-8)
            gen = torch.Generator("cuda").manual_seed(5)
            for _ in range(3):
                x = torch.randn(8, 20, device="cuda", generator=gen)
                opt.zero_grad()
                net(x).pow(2).mean().backward()
                opt.step()
            torch.cuda.synchronize()
            res[fuse] = ([p.detach().cpu().numpy() for p in net.parameters()],
                         [p.grad.cpu().numpy() for p in net.parameters()])
        finally:
            optim.FUSE_STEP = True
    for a, b in zip(res[True][0], res[False][0]):
        assert_bits_equal(a, b, "parameters after in-place steps")
    for a, b in zip(res[True][1], res[False][1]):
        assert_bits_equal(a, b, "p.grad holds the scaled update after step, as in the reference's in-place chain")


def test_adam_step_none_grads_and_multi_dtype():
    """Leaves without gradients pass through; fp32 and fp64 leaves in one call are grouped into two launches."""
    import torchopt_b200 as tb
    step = tb.AdamStep(step_size=-1e-2, eps_root=1e-8)
    p = [torch.randn(100, device="cuda"), torch.randn(50, device="cuda", dtype=torch.float64), torch.randn(3, device="cuda")]
    g = [torch.randn(100, device="cuda"), torch.randn(50, device="cuda", dtype=torch.float64), None]
    mu = [torch.zeros_like(t) for t in p]
    nu = [torch.zeros_like(t) for t in p]
    before = tb.abi.kernel_launches()
    new_p, new_mu, new_nu = step(p, g, mu, nu, [1, 1, 1])
    assert tb.abi.kernel_launches() - before == 2
    assert new_p[2] is p[2] and new_mu[2] is mu[2]
    for i in (0, 1):
        m, v, u = tb.AdamOp(eps_root=1e-8, inplace=False)(mu[i], nu[i], g[i], 1)
        assert torch.equal(new_p[i], p[i] + u * (-1e-2)) and torch.equal(new_mu[i], m) and torch.equal(new_nu[i], v)


# ---------------------------------------------------------------------------------------------------------------
# weight decay (L2 and decoupled) and maximize folded into the whole-step launch
# ---------------------------------------------------------------------------------------------------------------
RECIPES = [dict(weight_decay=1e-2, decoupled=False, maximize=False), dict(weight_decay=1e-2, decoupled=True, maximize=False),
           dict(weight_decay=0.0, decoupled=False, maximize=True), dict(weight_decay=3e-2, decoupled=False, maximize=True),
           dict(weight_decay=3e-2, decoupled=True, maximize=True)]


NO_DECAY = dict(weight_decay=0.0, decoupled=False, maximize=False)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.bfloat16])
@pytest.mark.parametrize("recipe", RECIPES + [NO_DECAY])
@pytest.mark.parametrize("detached", ["", "g", "g,mu,nu", "p"])
def test_adam_step_recipes_equal_torch_chain(dtype, recipe, detached):
    """AdamStep (one launch each way) vs the reference's chain evaluated with torch ops on the same GPU:
    g.neg(), g.add(p, alpha=wd), the Adam operator, u.add(p, alpha=wd), u.mul(-lr), p.add(u) -- outputs and all
    gradients (w.r.t. params, grads, moments) bit-identical.  ``detached`` names the inputs that do NOT require grad:
    with first-order (detached) gradients and L2 decay the parameter gradient still needs WD1 * dg2."""
    import torchopt_b200 as tb
    lr, er = 2e-3, 1e-8
    wd, dec, mx = recipe["weight_decay"], recipe["decoupled"], recipe["maximize"]
    gen = torch.Generator("cuda").manual_seed(3)
    sizes = [1, 9, 1000, 4097, 33_001]

    # bf16 parameters / gradients keep fp32 moments (the "bf16-param / fp32-state" variant)
    state_dtype = torch.float32 if dtype is torch.bfloat16 else dtype

    def leaves(scale, positive=False, dt=dtype):
        out = []
        for k in sizes:
            t = torch.randn(k, device="cuda", dtype=torch.float32, generator=gen) * scale
            out.append((t.abs() if positive else t).to(dt).requires_grad_(True))
        return out

    p, g = leaves(0.3), leaves(1e-2)
    mu, nu = leaves(1e-2, dt=state_dtype), leaves(1e-4, positive=True, dt=state_dtype)
    douts = [[torch.randn(k, device="cuda", dtype=torch.float32, generator=gen).to(dt) for k in sizes]
             for dt in (dtype, state_dtype, state_dtype)]
    with torch.no_grad():
        for t in g + mu:
            t[::7] = 0
    off = set(detached.split(",")) - {""}
    if "p" in off:
        p = [t.detach() for t in p]
    if "g" in off:
        g = [t.detach() for t in g]
    if "mu" in off:
        mu = [t.detach() for t in mu]
    if "nu" in off:
        nu = [t.detach() for t in nu]
    wrt = [(name, t) for name, ts in (("dp", p), ("dg", g), ("dmu", mu), ("dnu", nu)) for t in ts if t.requires_grad]

    def chain():
        op = tb.AdamOp(eps_root=er, inplace=False)
        outs = []
        for i in range(len(sizes)):
            gi = g[i].neg() if mx else g[i]
            if wd and not dec:
                gi = gi.add(p[i], alpha=wd)
            m2, v2, u = op(mu[i], nu[i], gi, 2)
            if wd and dec:
                u = u.add(p[i], alpha=wd)
            outs.append((p[i].add(u.mul(-lr)), m2, v2))
        return [list(x) for x in zip(*outs)]

    def fused():
        step = tb.AdamStep(eps_root=er, step_size=-lr, **recipe)
        return [list(x) for x in step(p, g, mu, nu, [2] * len(sizes))]

    res = []
    for fn in (chain, fused):
        new_p, new_mu, new_nu = fn()
        total = sum((a * d).sum() for outs, ds in zip((new_p, new_mu, new_nu), douts) for a, d in zip(outs, ds))
        grads = torch.autograd.grad(total, [t for _, t in wrt])
        torch.cuda.synchronize()
        res.append(([t.detach() for t in new_p + new_mu + new_nu], grads))
    for a, b in zip(res[0][0], res[1][0]):
        assert torch.equal(a, b), "forward outputs differ from the torch chain"
    for (name, _), a, b in zip(wrt, res[0][1], res[1][1]):
        assert torch.equal(a, b), f"{name} differs from the torch chain (detached: {detached!r})"


@pytest.mark.parametrize("cls_name,wd,mx", [("MetaAdam", 1e-2, False), ("MetaAdam", 1e-2, True), ("MetaAdamW", 1e-2, False),
                                            ("MetaAdamW", 5e-2, True)])
def test_meta_optimizers_with_weight_decay_fused_equals_chained(cls_name, wd, mx):
    import torchopt_b200 as tb
    from torchopt_b200 import optim
    x = torch.randn(16, 20, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    y = torch.randn(16, 7, device="cuda", generator=torch.Generator("cuda").manual_seed(2))
    out = {}
    for fuse in (True, False):
        optim.FUSE_STEP = fuse
        try:
            net = _mlp()
            init = list(net.parameters())
            opt = getattr(tb, cls_name)(net, lr=1e-2, weight_decay=wd, maximize=mx, eps_root=1e-8)
            for _ in range(3):
                opt.step(((net(x) - y) ** 2).mean())
            meta = torch.autograd.grad(((net(x) - y) ** 2).mean(), init)
            torch.cuda.synchronize()
            out[fuse] = ([p.detach() for p in opt.params()], meta)
        finally:
            optim.FUSE_STEP = True
    for a, b in zip(out[True][0] + list(out[True][1]), out[False][0] + list(out[False][1])):
        assert torch.equal(a, b)


@pytest.mark.parametrize("decoupled", [False, True])
@pytest.mark.parametrize("maximize", [False, True])
def test_inplace_optimizers_with_weight_decay(decoupled, maximize):
    """In-place Adam(weight_decay) / AdamW: fused launch == chained path bit-for-bit, and both track torch.optim
    (fp64, 25x default tolerances on parameter deltas, as the reference's tests/test_optim.py)."""
    import torchopt_b200 as tb
    from torchopt_b200 import optim
    cls, ref_cls = (tb.AdamW, torch.optim.AdamW) if decoupled else (tb.Adam, torch.optim.Adam)
    res = {}
    for fuse in (True, False, "torch"):
        optim.FUSE_STEP = fuse is True
        try:
            net = _mlp(torch.float64)
            base = [p.detach().clone() for p in net.parameters()]
            opt = (ref_cls(net.parameters(), lr=1e-3, weight_decay=1e-2, maximize=maximize) if fuse == "torch"
                   else cls(net.parameters(), lr=1e-3, weight_decay=1e-2, maximize=maximize))
            gen = torch.Generator("cuda").manual_seed(5)
            for _ in range(5):
                x = torch.randn(8, 20, device="cuda", dtype=torch.float64, generator=gen)
                opt.zero_grad()
                net(x).pow(2).mean().backward()
                opt.step()
            torch.cuda.synchronize()
            res[fuse] = [p.detach() - b for p, b in zip(net.parameters(), base)]
        finally:
            optim.FUSE_STEP = True
    for a, b, c in zip(res[True], res[False], res["torch"]):
        assert torch.equal(a, b)
        torch.testing.assert_close(a, c, rtol=2.5e-6, atol=2.5e-6)


def test_meta_adam_on_bf16_model_uses_fp32_moments():
    """A bf16 model through MetaAdam: moments are created in fp32, the whole-step launch handles the mixed family and
    equals the chained path (fused Adam op + torch's bf16 mul/add) bit-for-bit, parameters and meta-gradients."""
    import torchopt_b200 as tb
    from torchopt_b200 import optim
    x = torch.randn(16, 20, device="cuda", generator=torch.Generator("cuda").manual_seed(1)).bfloat16()
    y = torch.randn(16, 7, device="cuda", generator=torch.Generator("cuda").manual_seed(2)).bfloat16()
    out = {}
    for fuse in (True, False):
        optim.FUSE_STEP = fuse
        try:
            net = _mlp(torch.bfloat16)
            init = list(net.parameters())
            opt = tb.MetaAdam(net, lr=1e-2, eps_root=1e-8, weight_decay=1e-2)
            for _ in range(3):
                opt.step(((net(x) - y) ** 2).mean())
            adam_state = next(st for st in opt.state if isinstance(st, tb.ScaleByAdamState))
            assert all(m.dtype == torch.float32 for m in adam_state.mu + adam_state.nu)
            assert all(p.dtype == torch.bfloat16 for p in opt.params())
            meta = torch.autograd.grad(((net(x) - y) ** 2).mean(), init)
            torch.cuda.synchronize()
            out[fuse] = [p.detach() for p in opt.params()] + list(meta) + [m.detach() for m in adam_state.mu]
        finally:
            optim.FUSE_STEP = True
    for a, b in zip(out[True], out[False]):
        assert torch.equal(a, b)
