"""GPU edge cases: degenerate hyper-parameters / counts (inf and NaN must propagate exactly as in the reference
CPU build), fp64 and bf16 NULL patterns, CUDA-graph capture of the fused launches, non-default streams."""
from __future__ import annotations

import numpy as np
import pytest

from conftest import assert_bits_equal

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


@pytest.mark.parametrize("npdt", [np.float32, np.float64])
def test_degenerate_hyper_and_counts(c_oracle, npdt):
    """count = 0 makes the bias corrections infinite; b = 0 removes the moving average; eps = 0 with nu' = 0 divides
    by zero.  IEEE inf/NaN must come out in the same places as in the reference CPU arithmetic."""
    from oracle import adam_oracle as ao
    from torchopt_b200 import abi
    dt = abi.F32 if npdt == np.float32 else abi.F64
    n = 2051
    rng = np.random.default_rng(1)
    g = (rng.standard_normal(n) * 1e-2).astype(npdt)
    mu = (rng.standard_normal(n) * 1e-2).astype(npdt)
    nu = (rng.random(n) * 1e-4).astype(npdt)
    du = rng.standard_normal(n).astype(npdt)
    dme = (rng.standard_normal(n) * 1e-3).astype(npdt)
    dne = (rng.standard_normal(n) * 1e-3).astype(npdt)
    g[::5] = 0
    mu[::5] = 0
    nu[::5] = 0
    nu[3::7] = -1e-6          # negative second moment -> sqrt of a negative number -> NaN
    g[4::11] = np.finfo(npdt).max
    G, MU, NU, DU, DME, DNE = map(dev, (g, mu, nu, du, dme, dne))
    o = [torch.empty_like(G) for _ in range(6)]
    p = [t.data_ptr() for t in o]
    s = torch.cuda.current_stream().cuda_stream
    cases = [(0.9, 0.999, 1e-8, 0.0, 0), (0.0, 0.0, 0.0, 0.0, 1), (0.9, 0.999, 0.0, 0.0, 1),
             (0.5, 0.25, 1e-3, 1e-3, 1_000_000_007), (0.999999, 0.9999999, 1e-8, 1e-12, 2 ** 40),
             (0.9, 0.999, 1e-8, 1e-8, 1)]
    for b1, b2, eps, er, cnt in cases:
        w_mu, w_nu, w_u = ao.fused_forward(c_oracle, g, mu, nu, b1, b2, eps, er, cnt)
        abi.fused_forward(dt, [G.data_ptr()], [MU.data_ptr()], [NU.data_ptr()], [p[0]], [p[1]], [p[2]], [n], [cnt], b1, b2,
                          eps, er, s)
        for got, want, name in zip(o[:3], (w_mu, w_nu, w_u), ("mu'", "nu'", "u")):
            assert_bits_equal(host(got), want, f"fwd {name} case {(b1, b2, eps, er, cnt)}")
        want = ao.fused_backward(c_oracle, du, w_u, w_mu, g, dme, dne, b1, b2, er, cnt)
        abi.fused_backward(dt, [DU.data_ptr()], [p[2]], [p[0]], [G.data_ptr()], [DME.data_ptr()], [DNE.data_ptr()], [p[3]],
                           [p[4]], [p[5]], [n], [cnt], b1, b2, er, s)
        for got, wv, name in zip(o[3:], want, ("dg", "dmu", "dnu")):
            assert_bits_equal(host(got), wv, f"bwd {name} case {(b1, b2, eps, er, cnt)}")


def test_generic_patterns_fp64_and_bf16(c_oracle):
    from oracle import adam_oracle as ao
    from torchopt_b200 import abi
    n, er, cnt = 3001, 1e-8, 2
    rng = np.random.default_rng(9)
    s = torch.cuda.current_stream().cuda_stream
    # fp64
    g, mu, nu, du, dme, dne = ((rng.standard_normal(n) * sc).astype(np.float64) for sc in (1e-2, 1e-2, 1e-4, 1, 1e-3, 1e-3))
    nu = np.abs(nu)
    mu[::9] = 0
    g[::9] = 0
    w_mu, w_nu, w_u = ao.fused_forward(c_oracle, g, mu, nu, 0.9, 0.999, 1e-8, er, cnt)
    G, U, M, DU, DME, DNE = map(dev, (g, w_u, w_mu, du, dme, dne))
    for hdu, hme, hne in ((1, 0, 1), (0, 1, 1), (0, 1, 0), (1, 1, 0)):
        want = ao.fused_backward(c_oracle, du if hdu else None, w_u, w_mu, g, dme if hme else None, dne if hne else None,
                                 0.9, 0.999, er, cnt)
        o = [torch.empty_like(G) for _ in range(3)]
        abi.fused_backward(abi.F64, [DU.data_ptr() if hdu else None], [U.data_ptr()], [M.data_ptr()], [G.data_ptr()],
                           [DME.data_ptr() if hme else None], [DNE.data_ptr() if hne else None], [o[0].data_ptr()],
                           [o[1].data_ptr()], [o[2].data_ptr()], [n], [cnt], 0.9, 0.999, er, s)
        for k in range(3):
            if want[k] is None:
                assert (host(o[k]) == 0).all()
            else:
                assert_bits_equal(host(o[k]), want[k], f"fp64 generic out{k} {(hdu, hme, hne)}")
    # bf16-mixed, last-step pattern and in-place forward
    g32 = (rng.standard_normal(n) * 1e-2).astype(np.float32)
    Gb = torch.from_numpy(g32).cuda().to(torch.bfloat16)
    gup = Gb.float().cpu().numpy()
    mu32, nu32 = mu.astype(np.float32), nu.astype(np.float32)
    MU, NU = dev(mu32), dev(nu32)
    a, b, c = Gb.clone(), MU.clone(), NU.clone()
    abi.forward_inplace(abi.BF16_F32, a.data_ptr(), b.data_ptr(), c.data_ptr(), n, 0.9, 0.999, 1e-8, er, cnt, s)
    w_mu, w_nu, w_u = ao.fused_forward(c_oracle, gup, mu32, nu32, 0.9, 0.999, 1e-8, er, cnt)
    assert_bits_equal(host(b), w_mu, "bf16 forward_ mu")
    assert_bits_equal(host(c), w_nu, "bf16 forward_ nu")
    torch.cuda.synchronize()
    assert torch.equal(a.cpu().view(torch.int16), torch.from_numpy(w_u).to(torch.bfloat16).view(torch.int16))


def test_cuda_graph_capture_and_side_stream():
    """The leaf table travels in the kernel parameters, so the fused launches can be captured in a CUDA graph and
    replayed (launch-bound unrolls), and they honour the caller's current stream."""
    import torchopt_b200 as tb
    torch.manual_seed(0)
    sizes = [5, 64, 320, 4096, 36_864, 100_003]
    g = [torch.randn(k, device="cuda") * 1e-2 for k in sizes]
    mu0 = [torch.randn(k, device="cuda") * 1e-2 for k in sizes]
    nu0 = [torch.rand(k, device="cuda") * 1e-4 for k in sizes]

    def unroll():
        mu, nu, us = mu0, nu0, None
        for k in range(1, 6):
            mu, nu, us = tb.adam_op.fused_forward(g, mu, nu, 0.9, 0.999, 1e-8, 1e-8, [k] * len(sizes), False)
        return mu, nu, us

    eager = unroll()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        on_side = unroll()
    side.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        captured = unroll()
    before = tb.abi.kernel_launches()
    graph.replay()
    graph.replay()
    torch.cuda.synchronize()
    assert tb.abi.kernel_launches() == before, "replays do not go through the host launch path"
    for a, b, c in zip(eager, on_side, captured):
        for x, y, z in zip(a, b, c):
            assert torch.equal(x, y) and torch.equal(x, z)


@pytest.mark.gpu
def test_prepared_launch_handle_equals_per_call_arguments():
    """b200_adam_prepare_fused / _launch_prepared / _release_prepared: the argument block kept on the library side gives
    the same launches as passing every argument per call (two leaves, ragged sizes); invalid arguments give NULL / E_ARG."""
    import torch
    from torchopt_b200 import abi
    lib = abi.load()
    sizes = [1000, 70_001]
    gen = torch.Generator("cuda").manual_seed(4)
    t = {k: [torch.randn(n, device="cuda", generator=gen) * sc for n in sizes]
         for k, sc in (("g", 1e-2), ("mu", 1e-2), ("du", 1.0), ("dme", 1e-3), ("dne", 1e-3))}
    t["nu"] = [torch.rand(n, device="cuda", generator=gen) * 1e-4 for n in sizes]
    outs = [{k: [torch.full((n,), 7.0, device="cuda") for n in sizes] for k in ("mu2", "nu2", "u", "dg", "dmu", "dnu")}
            for _ in range(2)]
    P = lambda d, k: [x.data_ptr() for x in d[k]]   # noqa: E731
    s = torch.cuda.current_stream().cuda_stream
    preps = []
    for o in outs:
        preps.append(abi.PreparedFused(
            abi.F32, [P(t, "g"), P(t, "mu"), P(t, "nu"), P(o, "mu2"), P(o, "nu2"), P(o, "u")],
            [P(t, "du"), P(o, "u"), P(o, "mu2"), P(t, "g"), P(t, "dme"), P(t, "dne"), P(o, "dg"), P(o, "dmu"), P(o, "dnu")],
            sizes, [3, 9], 0.9, 0.999, 1e-8, 1e-8))
    assert preps[0].handle and preps[1].handle
    preps[1].lib.b200_adam_release_prepared(preps[1].handle)
    preps[1].handle = None                          # second instance: the per-call path
    for p in preps:
        assert p.forward(s) == 0 and p.backward(s) == 0
    torch.cuda.synchronize()
    for k in outs[0]:
        for a, b in zip(outs[0][k], outs[1][k]):
            assert torch.equal(a, b) and not (a == 7.0).all(), k
    assert lib.b200_adam_launch_prepared(None, 0, s) == -2
    assert not lib.b200_adam_prepare_fused(abi.F32, 0, *([None] * 12), None, None, 0.9, 0.999, 1e-8, 0.0)
