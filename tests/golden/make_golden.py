"""Generate tests/golden/*.npz from the REFERENCE ITSELF (run in the build container only).

The reference's tests hold no golden vectors for this path (forward pinned only
loosely against other optimizers, backward not at all), so the fixtures are outputs of the reference's
own compiled ``adam_op`` (oracle/_ref/_C.so, unmodified sources, built by oracle/build_ref.py) and of its
own Python ``AdamOp`` autograd dispatch (/root/reference/torchopt/accelerated_op/adam_op.py, imported by
file path) on seeded inputs.  Inputs come from numpy's PCG64 ``default_rng(seed)`` so the .npz files are
self-contained: they store inputs AND outputs.

    python tests/golden/make_golden.py        # needs /root/reference + oracle/_ref/_C.so

Files:
  adam_ops_{f32,f64}.npz   -- the 7 ops on n=773 random + 27 special-value elements, 6 (eps_root,count) cases
  adam_unroll_{f32,f64}.npz-- 5-step unrolled differentiable Adam through the reference AdamOp + autograd:
                              final params, meta-gradients w.r.t. initial params / per-step loss weights /
                              initial moments, for moment_requires_grad in {False, True}
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import adam_oracle as ao  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ADAM_OP_PY = "/root/reference/torchopt/accelerated_op/adam_op.py"

B1, B2, EPS = 0.9, 0.999, 1e-8
CASES = [(0.0, 1), (0.0, 2), (0.0, 5), (1e-8, 1), (1e-8, 3), (1e-3, 1000)]


def specials(dt):
    fi = np.finfo(dt)
    return np.array([0.0, -0.0, 1.0, -1.0, fi.tiny, -fi.tiny, fi.tiny / 8, fi.max, -fi.max, fi.eps,
                     np.inf, -np.inf, np.nan, 1e-20, -1e-20, 1e20, 3.0, 1 / 3, 1e-8, 65504.0,
                     2.0 ** -24, 0.1, 0.999, 123456.789, -0.5, 7e-5, 1e-30], dtype=dt)


def make_inputs(dt, seed=20240613, n=773):
    rng = np.random.default_rng(seed)
    sp = specials(dt)
    k = sp.size

    def cat(a, roll):
        return np.concatenate([a.astype(dt), np.roll(sp, roll)])

    g = cat(rng.standard_normal(n) * 1e-2, 0)
    mu = cat(rng.standard_normal(n) * 1e-2, 3)
    nu = cat(rng.random(n) * 1e-4, 7)
    du = cat(rng.standard_normal(n), 11)
    dmu = cat(rng.standard_normal(n) * 1e-3, 5)
    dnu = cat(rng.standard_normal(n) * 1e-3, 13)
    g[::97] = 0
    mu[::97] = 0  # -> new_mu == 0 exactly at those indices (zero branch of backward_updates)
    assert g.size == n + k
    return dict(g=g, mu=mu, nu=nu, du=du, dmu=dmu, dnu=dnu)


def gen_ops(ref, dt, tag):
    inp = make_inputs(dt)
    out = dict(inp)
    t = {k: torch.from_numpy(v.copy()) for k, v in inp.items()}
    out["forward_mu"] = ref.forward_mu(t["g"], t["mu"], B1).numpy()
    out["forward_nu"] = ref.forward_nu(t["g"], t["nu"], B2).numpy()
    bm = ref.backward_mu(t["dmu"], t["g"], t["mu"], B1)
    out["backward_mu.dg"], out["backward_mu.dmu"] = bm[0].numpy(), bm[1].numpy()
    bn = ref.backward_nu(t["dnu"], t["g"], t["nu"], B2)
    out["backward_nu.dg"], out["backward_nu.dnu"] = bn[0].numpy(), bn[1].numpy()
    new_mu = torch.from_numpy(out["forward_mu"].copy())
    new_nu = torch.from_numpy(out["forward_nu"].copy())
    for ci, (er, cnt) in enumerate(CASES):
        u = ref.forward_updates(new_mu, new_nu, B1, B2, EPS, er, cnt)
        out[f"case{ci}.forward_updates"] = u.numpy()
        a, b, c = t["g"].clone(), t["mu"].clone(), t["nu"].clone()
        r = ref.forward_(a, b, c, B1, B2, EPS, er, cnt)
        assert r[0].data_ptr() == a.data_ptr() and r[1].data_ptr() == b.data_ptr()
        out[f"case{ci}.forward_.updates"], out[f"case{ci}.forward_.mu"], out[f"case{ci}.forward_.nu"] = (
            a.numpy(), b.numpy(), c.numpy())
        bu = ref.backward_updates(t["du"], u, new_mu, new_nu, B1, B2, er, cnt)
        out[f"case{ci}.backward_updates.dmu"], out[f"case{ci}.backward_updates.dnu"] = (
            bu[0].numpy(), bu[1].numpy())
    out["cases"] = np.array(CASES, dtype=np.float64)
    out["hyper"] = np.array([B1, B2, EPS])
    np.savez(os.path.join(HERE, f"adam_ops_{tag}.npz"), **out)
    print("wrote", f"adam_ops_{tag}.npz", {k: v.shape for k, v in list(out.items())[:3]})


def load_ref_adamop(ref_C):
    """Import the reference's accelerated_op/adam_op.py by path, bound to the compiled reference _C."""
    pkg = types.ModuleType("torchopt")
    pkg.__path__ = []
    sys.modules["torchopt"] = pkg
    sys.modules["torchopt._C"] = ref_C
    sys.modules["torchopt._C.adam_op"] = ref_C.adam_op
    spec = importlib.util.spec_from_file_location("torchopt.accelerated_op.adam_op", REF_ADAM_OP_PY)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.AdamOp


def unroll_inputs(dt, n=515, steps=5, seed=777):
    rng = np.random.default_rng(seed)
    p0 = (rng.standard_normal(n) * 0.5).astype(dt)
    p0[::64] = 0  # grads exactly 0 at every step -> new_mu == 0 branch in every backward
    w = (rng.random((steps, n)) + 0.5).astype(dt)
    tgt = rng.standard_normal(n).astype(dt)
    mu0 = (rng.standard_normal(n) * 1e-3).astype(dt)
    mu0[::64] = 0
    nu0 = (rng.random(n) * 1e-5).astype(dt)
    return p0, w, tgt, mu0, nu0


def run_unroll(AdamOp, p0, w, tgt, mu0, nu0, *, moment_requires_grad, eps_root, lr=1e-2):
    """5 inner steps: g_k = w_k * p_k (differentiable), Adam update, p_{k+1} = p_k - lr*u; outer = <p_K, tgt>^2-ish."""
    op = AdamOp(b1=B1, b2=B2, eps=EPS, eps_root=eps_root, inplace=False)
    p = torch.from_numpy(p0.copy()).requires_grad_(True)
    wt = torch.from_numpy(w.copy()).requires_grad_(True)
    mu = torch.from_numpy(mu0.copy()).requires_grad_(moment_requires_grad)
    nu = torch.from_numpy(nu0.copy()).requires_grad_(moment_requires_grad)
    t = torch.from_numpy(tgt.copy())
    leaves = [p, wt] + ([mu, nu] if moment_requires_grad else [])
    cur, cmu, cnu = p, mu, nu
    for k in range(w.shape[0]):
        g = wt[k] * cur
        cmu, cnu, u = op(cmu, cnu, g, k + 1)
        cur = cur + (-lr) * u
    outer = ((cur - t) ** 2).sum() + (cmu * t).sum() + (cnu * t * t).sum()
    grads = torch.autograd.grad(outer, leaves)
    res = {"p_final": cur.detach().numpy(), "mu_final": cmu.detach().numpy(),
           "nu_final": cnu.detach().numpy(), "outer": outer.detach().numpy(),
           "d_p0": grads[0].numpy(), "d_w": grads[1].numpy()}
    if moment_requires_grad:
        res["d_mu0"], res["d_nu0"] = grads[2].numpy(), grads[3].numpy()
    return res


def gen_unroll(AdamOp, dt, tag):
    p0, w, tgt, mu0, nu0 = unroll_inputs(dt)
    out = dict(p0=p0, w=w, tgt=tgt, mu0=mu0, nu0=nu0, lr=np.array(1e-2))
    for mrg in (False, True):
        for er in (0.0, 1e-8):
            res = run_unroll(AdamOp, p0, w, tgt, mu0, nu0, moment_requires_grad=mrg, eps_root=er)
            for k, v in res.items():
                out[f"mrg{int(mrg)}.er{er:g}.{k}"] = v
    np.savez(os.path.join(HERE, f"adam_unroll_{tag}.npz"), **out)
    print("wrote", f"adam_unroll_{tag}.npz")


def main():
    ref_C = ao.load_ref()
    if ref_C is None:
        raise SystemExit("oracle/_ref/_C.so missing: run python oracle/build_ref.py first")
    torch.set_num_threads(1)
    for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
        gen_ops(ref_C.adam_op, dt, tag)
    AdamOp = load_ref_adamop(ref_C)
    for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
        gen_unroll(AdamOp, dt, tag)


if __name__ == "__main__":
    main()
