"""Parity of the host-buffer pipeline (`adam_op.host_step` -> `b200_adam_host_step_f32`, the call bench.py times as
`e2e`): host arrays in, host arrays out, bit-identical to the CPU oracle for every slicing of the arrays -- one slice,
several slices with a ragged last one, many slices rotating through the staging slots."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from conftest import assert_bits_equal

gpu = pytest.mark.gpu
B1, B2, EPS = 0.9, 0.999, 1e-8


def _inputs(n, seed):
    rng = np.random.default_rng(seed)
    g = (rng.standard_normal(n) * 1e-2).astype(np.float32)
    mu = (rng.standard_normal(n) * 1e-2).astype(np.float32)
    nu = (rng.random(n) * 1e-4).astype(np.float32)
    du = rng.standard_normal(n).astype(np.float32)
    dme = (rng.standard_normal(n) * 1e-3).astype(np.float32)
    dne = (rng.standard_normal(n) * 1e-3).astype(np.float32)
    g[::1024] = 0
    mu[::1024] = 0          # mu' == 0 exactly: the zero branch of backward_updates
    return g, mu, nu, du, dme, dne


@gpu
@pytest.mark.parametrize("n,slice_elems", [
    (1, 0), (7, 8), (1000, 0),                 # shorter than one slice
    (100_003, 32_768),                          # 4 slices, ragged last one
    (1_000_003, 65_536),                        # 16 slices: every staging slot is reused several times
    (4 * 65_536, 65_536),                       # exact multiple of the slice length
    (3_000_001, 1 << 20),
])
@pytest.mark.parametrize("eps_root,count", [(0.0, 1), (1e-8, 7)])
def test_host_step_equals_oracle(c_oracle, n, slice_elems, eps_root, count):
    import torchopt_b200 as tb
    from oracle import adam_oracle as ao
    arrs = _inputs(n, seed=n % 9973)
    pinned = [torch.from_numpy(a).pin_memory() for a in arrs]
    out = [torch.full((n,), float("nan")).pin_memory() for _ in range(6)]
    res = tb.adam_op.host_step(*pinned, out, B1, B2, EPS, eps_root, count, slice_elems)
    assert all(r.data_ptr() == o.data_ptr() for r, o in zip(res, out)), "results are written into the caller's arrays"
    g, mu, nu, du, dme, dne = arrs
    w_mu, w_nu, w_u = ao.fused_forward(c_oracle, g, mu, nu, B1, B2, EPS, eps_root, count)
    w_dg, w_dmu, w_dnu = ao.fused_backward(c_oracle, du, w_u, w_mu, g, dme, dne, B1, B2, eps_root, count)
    for name, got, want in zip(("mu'", "nu'", "u", "dg", "dmu", "dnu"), out, (w_mu, w_nu, w_u, w_dg, w_dmu, w_dnu)):
        assert_bits_equal(got.numpy(), want, f"host_step {name} (n={n}, slice={slice_elems})")
    tb._C.host_release()


@gpu
def test_host_step_pageable_memory_and_argument_errors():
    """Pageable (unpinned) host arrays are refused unless asked for (the driver stages them and the three streams of the
    pipeline serialise), and then give the same results; bad arguments raise."""
    import torchopt_b200 as tb
    n = 50_000
    arrs = [torch.from_numpy(a) for a in _inputs(n, 1)]
    out = [torch.empty(n) for _ in range(6)]
    with pytest.raises(RuntimeError, match="page-locked"):
        tb.adam_op.host_step(*arrs, out, B1, B2, EPS, 0.0, 3, 16_384)
    tb.adam_op.host_step(*arrs, out, B1, B2, EPS, 0.0, 3, 16_384, allow_pageable=True)
    pinned_out = [torch.empty(n).pin_memory() for _ in range(6)]
    tb.adam_op.host_step(*[a.pin_memory() for a in arrs], pinned_out, B1, B2, EPS, 0.0, 3, 0)
    for a, b in zip(out, pinned_out):
        assert torch.equal(a, b)
    with pytest.raises(RuntimeError):
        tb.adam_op.host_step(*arrs, [torch.empty(n - 1) for _ in range(6)], B1, B2, EPS, 0.0, 3, 0, allow_pageable=True)
    with pytest.raises(RuntimeError):
        tb.adam_op.host_step(*[a.double() for a in arrs], out, B1, B2, EPS, 0.0, 3, 0, allow_pageable=True)
    tb._C.host_release()
