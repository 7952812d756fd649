"""Pins the oracle: C and NumPy restatements vs the golden vectors produced by the reference itself
(tests/golden/make_golden.py) and, when oracle/_ref/_C.so is present, vs the reference live."""
from __future__ import annotations

import numpy as np
import pytest

from conftest import assert_bits_equal

TAGS = {"f32": np.float32, "f64": np.float64}


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("which", ["c", "np"])
def test_oracle_matches_golden_ops(golden, c_oracle, np_oracle, tag, which):
    o = c_oracle if which == "c" else np_oracle
    z = golden(f"adam_ops_{tag}.npz")
    b1, b2, eps = z["hyper"]
    g, mu, nu, du, dmu, dnu = (z[k] for k in ("g", "mu", "nu", "du", "dmu", "dnu"))
    assert g.dtype == TAGS[tag]
    new_mu, new_nu = o.forward_mu(g, mu, b1), o.forward_nu(g, nu, b2)
    assert_bits_equal(new_mu, z["forward_mu"], "forward_mu")
    assert_bits_equal(new_nu, z["forward_nu"], "forward_nu")
    a, b = o.backward_mu(dmu, b1)
    assert_bits_equal(a, z["backward_mu.dg"], "backward_mu.dg")
    assert_bits_equal(b, z["backward_mu.dmu"], "backward_mu.dmu")
    a, b = o.backward_nu(dnu, g, b2)
    assert_bits_equal(a, z["backward_nu.dg"], "backward_nu.dg")
    assert_bits_equal(b, z["backward_nu.dnu"], "backward_nu.dnu")
    for ci, (er, cnt) in enumerate(z["cases"]):
        cnt = int(cnt)
        u = o.forward_updates(new_mu, new_nu, b1, b2, eps, er, cnt)
        assert_bits_equal(u, z[f"case{ci}.forward_updates"], f"forward_updates case{ci}")
        if which == "c":
            a, b, c = g.copy(), mu.copy(), nu.copy()
            o.forward_(a, b, c, b1, b2, eps, er, cnt)
        else:
            a, b, c = o.forward_(g, mu, nu, b1, b2, eps, er, cnt)
        assert_bits_equal(a, z[f"case{ci}.forward_.updates"], "forward_.updates")
        assert_bits_equal(b, z[f"case{ci}.forward_.mu"], "forward_.mu")
        assert_bits_equal(c, z[f"case{ci}.forward_.nu"], "forward_.nu")
        a, b = o.backward_updates(du, u, new_mu, b1, b2, er, cnt)
        assert_bits_equal(a, z[f"case{ci}.backward_updates.dmu"], f"backward_updates.dmu case{ci}")
        assert_bits_equal(b, z[f"case{ci}.backward_updates.dnu"], f"backward_updates.dnu case{ci}")


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_fused_composition_matches_golden_unroll(golden, c_oracle, tag):
    """The fused fwd/bwd composition used as the checker of the fused CUDA kernels reproduces the reference's
    AdamOp + autograd 5-step unroll bit-for-bit (forward values) -- hand-rolled reverse sweep for the grads."""
    from oracle import adam_oracle as ao
    z = golden(f"adam_unroll_{tag}.npz")
    dt = TAGS[tag]
    p0, w, tgt, mu0, nu0 = z["p0"], z["w"], z["tgt"], z["mu0"], z["nu0"]
    lr = dt(-float(z["lr"]))
    for mrg in (0, 1):
        for er in (0.0, 1e-8):
            key = f"mrg{mrg}.er{er:g}."
            p, mu, nu = p0, mu0, nu0
            saved = []
            for k in range(w.shape[0]):
                g = w[k] * p
                new_mu, new_nu, u = ao.fused_forward(c_oracle, g, mu, nu, 0.9, 0.999, 1e-8, er, k + 1)
                saved.append((p, g, u, new_mu))
                p, mu, nu = p + lr * u, new_mu, new_nu
            assert_bits_equal(p, z[key + "p_final"], "p_final")
            assert_bits_equal(mu, z[key + "mu_final"], "mu_final")
            assert_bits_equal(nu, z[key + "nu_final"], "nu_final")
            # reverse sweep: outer = sum((p-t)^2) + sum(mu*t) + sum(nu*t*t)
            dp = dt(2) * (p - tgt)
            dmu_ext, dnu_ext = tgt.copy(), tgt * tgt
            d_w = np.zeros_like(w)
            for k in reversed(range(w.shape[0])):
                pk, g, u, new_mu = saved[k]
                du = dp * lr
                dg, dmu_ext, dnu_ext = ao.fused_backward(c_oracle, du, u, new_mu, g, dmu_ext, dnu_ext,
                                                         0.9, 0.999, er, k + 1)
                d_w[k] = dg * pk
                dp = dp + dg * w[k]
            np.testing.assert_allclose(dp, z[key + "d_p0"], rtol=1e-5 if tag == "f32" else 1e-12, atol=0)
            np.testing.assert_allclose(d_w, z[key + "d_w"], rtol=1e-5 if tag == "f32" else 1e-12, atol=0)
            if mrg:
                assert_bits_equal(dmu_ext, z[key + "d_mu0"], "d_mu0")
                assert_bits_equal(dnu_ext, z[key + "d_nu0"], "d_nu0")


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_oracles_match_reference_live(c_oracle, np_oracle, dt):
    from oracle import adam_oracle as ao
    ref = ao.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref/_C.so not built (reference sources absent)")
    import torch
    ref = ref.adam_op
    rng = np.random.default_rng(5)
    for n in (0, 1, 3, 999, 1000, 1001, 40_003):
        g = (rng.standard_normal(n) * 1e-2).astype(dt)
        mu = (rng.standard_normal(n) * 1e-2).astype(dt)
        nu = (rng.random(n) * 1e-4).astype(dt)
        du = rng.standard_normal(n).astype(dt)
        mu[::7] = 0
        g[::7] = 0
        T = torch.from_numpy
        for er, cnt in ((0.0, 1), (1e-8, 4)):
            rm = ref.forward_mu(T(g), T(mu), 0.9).numpy()
            rn = ref.forward_nu(T(g), T(nu), 0.999).numpy()
            ru = ref.forward_updates(T(rm), T(rn), 0.9, 0.999, 1e-8, er, cnt).numpy()
            rdm, rdn = (t.numpy() for t in ref.backward_updates(T(du), T(ru), T(rm), T(rn), 0.9, 0.999, er, cnt))
            for o in (c_oracle, np_oracle):
                assert_bits_equal(o.forward_mu(g, mu, 0.9), rm, "forward_mu")
                assert_bits_equal(o.forward_nu(g, nu, 0.999), rn, "forward_nu")
                assert_bits_equal(o.forward_updates(rm, rn, 0.9, 0.999, 1e-8, er, cnt), ru, "forward_updates")
                a, b = o.backward_updates(du, ru, rm, 0.9, 0.999, er, cnt)
                assert_bits_equal(a, rdm, "backward_updates.dmu")
                assert_bits_equal(b, rdn, "backward_updates.dnu")


def test_host_scalars_match_c(c_oracle):
    """Scalar derivation (pow in double, one rounding) is identical between the numpy and the C oracle."""
    import ctypes
    from oracle import adam_oracle as ao
    lib = c_oracle.lib
    for f in ("oracle_adam_inv_one_minus_pow", "oracle_adam_one_minus_pow", "oracle_adam_inv_one_minus_pow_eps"):
        getattr(lib, f).restype = ctypes.c_double
    for b1, b2 in ((0.9, 0.999), (0.5, 0.75), (0.0, 0.0), (0.99, 0.9999)):
        for cnt in (1, 2, 3, 10, 1000, 10 ** 6):
            s = ao.host_scalars(np.float64, b1, b2, 1e-8, 1e-8, cnt)
            assert s["C1"] == lib.oracle_adam_inv_one_minus_pow(ctypes.c_double(b1), ctypes.c_uint64(cnt))
            assert s["C2"] == lib.oracle_adam_inv_one_minus_pow(ctypes.c_double(b2), ctypes.c_uint64(cnt))
            assert s["O1"] == lib.oracle_adam_one_minus_pow(ctypes.c_double(b1), ctypes.c_uint64(cnt))
            assert s["C2E"] == lib.oracle_adam_inv_one_minus_pow_eps(
                ctypes.c_double(b2), ctypes.c_double(1e-8), ctypes.c_uint64(cnt))


@pytest.mark.parametrize("tag,dt", [("f32", np.float32), ("f64", np.float64)])
def test_ref_unroll_reproduces_golden(golden, tag, dt):
    """oracle/ref_unroll.py (the restated Python glue of the reference's AdamOp over its compiled CPU kernels) reproduces
    the golden 5-step unroll -- values AND meta-gradients -- that tests/golden/make_golden.py recorded with the
    reference's own accelerated_op/adam_op.py.  This pins the harness bench.py uses for BASELINE config [0]."""
    import torch
    from oracle import adam_oracle as ao
    from oracle.ref_unroll import make_ref_adam_op
    ref = ao.load_ref()
    if ref is None:
        pytest.skip("oracle/_ref/_C.so not built")
    torch.set_num_threads(1)
    z = golden(f"adam_unroll_{tag}.npz")
    AdamOp = make_ref_adam_op(ref.adam_op)
    lr = float(z["lr"])
    for mrg in (False, True):
        for er in (0.0, 1e-8):
            op = AdamOp(b1=0.9, b2=0.999, eps=1e-8, eps_root=er, inplace=False)
            p = torch.from_numpy(z["p0"].copy()).requires_grad_(True)
            wt = torch.from_numpy(z["w"].copy()).requires_grad_(True)
            mu = torch.from_numpy(z["mu0"].copy()).requires_grad_(mrg)
            nu = torch.from_numpy(z["nu0"].copy()).requires_grad_(mrg)
            t = torch.from_numpy(z["tgt"].copy())
            cur, cmu, cnu = p, mu, nu
            for k in range(z["w"].shape[0]):
                cmu, cnu, u = op(cmu, cnu, wt[k] * cur, k + 1)
                cur = cur + (-lr) * u
            outer = ((cur - t) ** 2).sum() + (cmu * t).sum() + (cnu * t * t).sum()
            grads = torch.autograd.grad(outer, [p, wt] + ([mu, nu] if mrg else []))
            key = f"mrg{int(mrg)}.er{er:g}."
            assert_bits_equal(cur.detach().numpy(), z[key + "p_final"], key + "p_final")
            assert_bits_equal(cmu.detach().numpy(), z[key + "mu_final"], key + "mu_final")
            assert_bits_equal(grads[0].numpy(), z[key + "d_p0"], key + "d_p0")
            assert_bits_equal(grads[1].numpy(), z[key + "d_w"], key + "d_w")
            if mrg:
                assert_bits_equal(grads[2].numpy(), z[key + "d_mu0"], key + "d_mu0")
                assert_bits_equal(grads[3].numpy(), z[key + "d_nu0"], key + "d_nu0")
