"""Numerical parity with the reference's own CUDA build (BASELINE north star: "results must match the reference's
adam_op, CPU and CUDA builds, <= 2 ulp per element per step, bounded drift over a K-step unroll").

`oracle/_ref/cuda/_C.so` is the reference's UNMODIFIED src/adam_op/adam_op_impl_cuda.cu compiled for sm_100 by
`python oracle/build_ref.py --cuda` (4 minutes; not part of build()).  The test is skipped when that file is absent.

The CPU build is the bit-exact contract (tests/test_ops_gpu.py).  The CUDA build may differ from it -- and so from this
library -- in two ways:
  * nvcc contracts `a*b + c` into one FMA, i.e. one rounding instead of two: up to 1 ulp of the larger addend per sum,
    so the tolerance is stated in ulps of max(|result|, |largest addend|), not of the (possibly cancelled) result;
  * defects: elements of a ragged tail (n mod 1024 in {1,2,3}) are never written, and a `mu' == 0` element makes its
    thread `return` instead of `continue`.  The inputs here avoid both (n multiple of 1024, no exact zeros).
"""
from __future__ import annotations

import importlib.util
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CUDA = os.path.join(ROOT, "oracle", "_ref", "cuda", "_C.so")
B1, B2, EPS = 0.9, 0.999, 1e-8
ULP = float(np.finfo(np.float32).eps)       # 2^-23: spacing of floats in [1, 2)
MAX_ULPS = 2.0


@pytest.fixture(scope="module")
def ref():
    if not os.path.isfile(REF_CUDA):
        pytest.skip("oracle/_ref/cuda/_C.so not built (python oracle/build_ref.py --cuda)")
    spec = importlib.util.spec_from_file_location("_C", REF_CUDA)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.adam_op


def ulps(got, want, *scales):
    """Largest |got - want| in units of ulp(max(|want|, |scales...|)) (ulp of x taken as 2^-23 * |x|, an upper bound
    of the true spacing by at most 2x, i.e. the check is at least as strict as stated)."""
    ref_mag = want.abs()
    for s in scales:
        ref_mag = torch.maximum(ref_mag, s.abs())
    err = (got - want).abs() / (ref_mag * ULP).clamp_min(1e-45)
    return float(err.max())


def u_scale(g, mu_prev, new_nu, eps_root, count):
    """Magnitude u would have if mu' = b1*mu + (1-b1)*g had not cancelled: ulps of THIS are the right unit for
    differences in u that stem from the <= 1 ulp-of-the-larger-addend differences in mu' (SURVEY.md section 8c)."""
    c1 = 1.0 / (1.0 - B1 ** count)
    c2 = 1.0 / (1.0 - B2 ** count)
    addend = torch.maximum((B1 * mu_prev).abs(), ((1 - B1) * g).abs())
    return addend * c1 / ((new_nu * c2 + eps_root).sqrt() + EPS)


def inputs(n, seed):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    def r(scale):
        return torch.randn(n, device="cuda", generator=gen) * scale
    g, mu, du, dme, dne = r(1e-2), r(1e-2), r(1.0), r(1e-3), r(1e-3)
    nu = torch.rand(n, device="cuda", generator=gen) * 1e-4
    return g, mu, nu, du, dme, dne


@pytest.mark.parametrize("n", [1024, 1 << 20, 3 * 1024 * 1000])
@pytest.mark.parametrize("eps_root,count", [(0.0, 1), (1e-8, 4)])
def test_seven_operators_vs_reference_cuda(ref, n, eps_root, count):
    import torchopt_b200 as tb
    ours = tb.adam_op
    g, mu, nu, du, _, _ = inputs(n, n % 1000 + count)
    report = {}

    new_mu_r, new_mu = ref.forward_mu(g, mu, B1), ours.forward_mu(g, mu, B1)
    report["forward_mu"] = ulps(new_mu, new_mu_r, B1 * mu, (1 - B1) * g)
    new_nu_r, new_nu = ref.forward_nu(g, nu, B2), ours.forward_nu(g, nu, B2)
    report["forward_nu"] = ulps(new_nu, new_nu_r, B2 * nu, (1 - B2) * g * g)
    # same inputs for both (the reference's own mu', nu'), so that only this operator's rounding is compared
    u_r = ref.forward_updates(new_mu_r, new_nu_r, B1, B2, EPS, eps_root, count)
    u = ours.forward_updates(new_mu_r, new_nu_r, B1, B2, EPS, eps_root, count)
    report["forward_updates"] = ulps(u, u_r)
    a, b = ref.backward_mu(du, g, mu, B1), ours.backward_mu(du, g, mu, B1)
    report["backward_mu"] = max(ulps(x, y) for x, y in zip(b, a))
    a, b = ref.backward_nu(du, g, nu, B2), ours.backward_nu(du, g, nu, B2)
    report["backward_nu"] = max(ulps(x, y) for x, y in zip(b, a))
    a = ref.backward_updates(du, u_r, new_mu_r, new_nu_r, B1, B2, eps_root, count)
    b = ours.backward_updates(du, u_r, new_mu_r, new_nu_r, B1, B2, eps_root, count)
    report["backward_updates"] = max(ulps(x, y) for x, y in zip(b, a))
    # in-place fused forward
    gr, mr, nr = g.clone(), mu.clone(), nu.clone()
    go, mo, no = g.clone(), mu.clone(), nu.clone()
    ref.forward_(gr, mr, nr, B1, B2, EPS, eps_root, count)
    ours.forward_(go, mo, no, B1, B2, EPS, eps_root, count)
    report["forward_.mu"] = ulps(mo, mr, B1 * mu, (1 - B1) * g)
    report["forward_.nu"] = ulps(no, nr, B2 * nu, (1 - B2) * g * g)
    # u is computed from each side's own mu', nu' here: their <= 1 ulp differences propagate through one multiply, one
    # sqrt and one division on top of the operator's own <= 2 ulp, hence 2 extra ulps, in units of the un-cancelled u
    report["forward_.updates"] = ulps(go, gr, u_scale(g, mu, nr, eps_root, count))
    torch.cuda.synchronize()
    print("ulps vs reference CUDA:", {k: round(v, 3) for k, v in report.items()})
    bad = {k: v for k, v in report.items() if v > MAX_ULPS + (2.0 if k == "forward_.updates" else 0.0)}
    assert not bad, f"beyond {MAX_ULPS} ulp vs the reference CUDA build: {bad} (all: {report})"


def test_fused_unroll_drift_vs_reference_cuda(ref):
    """K = 5 unrolled steps driven by the same gradient sequence: this library's fused forward vs the reference CUDA
    operators chained the way AdamOp chains them.  Per-step differences (<= 1 ulp of the larger addend) must not
    amplify: moments stay within 2 ulp per step taken, updates within a few ulp."""
    import torchopt_b200 as tb
    n, K = 1 << 20, 5
    g0, mu, nu, _, _, _ = inputs(n, 77)
    mu_r, nu_r, mu_o, nu_o = mu.clone(), nu.clone(), mu.clone(), nu.clone()
    worst = {"mu": 0.0, "nu": 0.0, "u": 0.0}
    for k in range(1, K + 1):
        g = g0 * (1.0 + 0.1 * k)
        mu_r2 = ref.forward_mu(g, mu_r, B1)
        nu_r2 = ref.forward_nu(g, nu_r, B2)
        u_r = ref.forward_updates(mu_r2, nu_r2, B1, B2, EPS, 1e-8, k)
        (mu_o2,), (nu_o2,), (u_o,) = tb.adam_op.fused_forward([g], [mu_o], [nu_o], B1, B2, EPS, 1e-8, [k], False)
        worst["mu"] = max(worst["mu"], ulps(mu_o2, mu_r2, B1 * mu_r, (1 - B1) * g) / k)
        worst["nu"] = max(worst["nu"], ulps(nu_o2, nu_r2, B2 * nu_r, (1 - B2) * g * g) / k)
        worst["u"] = max(worst["u"], ulps(u_o, u_r, u_scale(g, mu_r, nu_r2, 1e-8, k)) / k)
        mu_r, nu_r, mu_o, nu_o = mu_r2, nu_r2, mu_o2, nu_o2
    torch.cuda.synchronize()
    print("worst ulps per step taken, 5-step unroll vs reference CUDA:", {k: round(v, 3) for k, v in worst.items()})
    # per step taken: moments within 2 ulp of the larger addend, updates within 4 ulp of the un-cancelled update
    assert worst["mu"] <= MAX_ULPS and worst["nu"] <= MAX_ULPS and worst["u"] <= 2 * MAX_ULPS, worst
