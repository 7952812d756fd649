"""Randomised A/B of the default state layout (flat moment buffers, ``FlatStepNode``) against the per-leaf reference
layout (``optim.FLAT_STATE = False``): random leaf sets (sizes incl. 0 and 1, fp32 / fp64 mixes), random recipes, random
per-step None patterns (incl. steps where nothing has a gradient), state reads / stop_gradient / state round trips at random
points, outer losses that use random subsets of the new parameters and of the moments.  Parameters, moments, counts and
meta-gradients must be bit-identical -- the two layouts run the same kernels on the same values."""
from __future__ import annotations

import os
import random

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

SIZE_POOL = [0, 1, 2, 7, 8, 9, 31, 33, 100, 1000, 4095, 4097, 20_001, 70_003]


def _scenario(seed):
    rnd = random.Random(seed)
    num = rnd.randint(1, 10)
    leaves = [(rnd.choice(SIZE_POOL), rnd.choice([torch.float32, torch.float32, torch.float64])) for _ in range(num)]
    steps = rnd.randint(1, 4)
    plan = []
    for _ in range(steps):
        mask = [rnd.random() < 0.8 for _ in range(num)]
        if rnd.random() < 0.1:
            mask = [False] * num
        plan.append(dict(mask=mask, action=rnd.choice(["none", "none", "read", "stop", "roundtrip", "roundtrip_copy"]),
                         detach_g=rnd.random() < 0.2))
    recipe = rnd.choice([dict(), dict(weight_decay=1e-2), dict(weight_decay=3e-2, maximize=True), dict(maximize=True)])
    cls = rnd.choice(["FuncAdam", "FuncAdamW"])
    if cls == "FuncAdamW" and "weight_decay" not in recipe:
        recipe = dict(recipe, weight_decay=1e-2)
    return dict(leaves=leaves, plan=plan, recipe=recipe, cls=cls, mrg=rnd.random() < 0.5,
                use_p=[rnd.random() < 0.8 for _ in range(num)], use_mu=[rnd.random() < 0.4 for _ in range(num)],
                eps_root=rnd.choice([0.0, 1e-8]), lr=rnd.choice([1e-2, 3e-3]))


def _run(sc, seed):
    import torchopt_b200 as tb
    gen = torch.Generator("cuda").manual_seed(seed)
    params = [(torch.randn(k, device="cuda", generator=gen) * 0.3).to(dt).requires_grad_(True) for k, dt in sc["leaves"]]
    ws = [[(torch.rand(k, device="cuda", generator=gen) + 0.5).to(dt) for k, dt in sc["leaves"]] for _ in sc["plan"]]
    opt = getattr(tb, sc["cls"])(sc["lr"], eps_root=sc["eps_root"], moment_requires_grad=sc["mrg"], **sc["recipe"])
    cur = list(params)
    for step, w in zip(sc["plan"], ws):
        grads = []
        for keep, wi, p in zip(step["mask"], w, cur):
            g = wi * p if keep else None
            grads.append(g.detach() if (g is not None and step["detach_g"]) else g)
        cur = opt.apply_gradients(grads, cur)
        if step["action"] == "read":
            _ = opt.state
        elif step["action"] == "stop":
            tb.stop_gradient(opt)
        elif step["action"] == "roundtrip" and opt.state is not None:
            opt.load_state_dict(tb.extract_state_dict(opt))
        elif step["action"] == "roundtrip_copy" and opt.state is not None:
            opt.load_state_dict(tb.extract_state_dict(opt, by="copy"))
    state = opt.state
    adam = None if state is None else next(s for s in state if isinstance(s, tb.ScaleByAdamState))
    terms = [(q.double() * q.double()).sum() for q, use in zip(cur, sc["use_p"]) if use and q.requires_grad]
    if adam is not None:
        terms += [(m.double() ** 2).sum() * 3.0 for m, use in zip(adam.mu, sc["use_mu"]) if use and m.requires_grad]
    meta = [None] * len(params)
    if terms:
        meta = torch.autograd.grad(torch.stack(terms).sum(), params, allow_unused=True)
    torch.cuda.synchronize()
    mu = [] if adam is None else [m.detach().clone() for m in adam.mu]
    nu = [] if adam is None else [v.detach().clone() for v in adam.nu]
    counts = [] if adam is None else [int(c) for c in adam.count]
    return [q.detach().clone() for q in cur], list(meta), mu, nu, counts


@pytest.mark.parametrize("seed", range(int(os.environ.get("B200_FUZZ_SEEDS", "40"))))
def test_flat_state_equals_per_leaf_state_on_random_scenarios(seed):
    from torchopt_b200 import optim
    sc = _scenario(seed)
    flat = _run(sc, seed)
    optim.FLAT_STATE = False
    try:
        leaf = _run(sc, seed)
    finally:
        optim.FLAT_STATE = True
    assert flat[4] == leaf[4], f"step counts differ: {flat[4]} vs {leaf[4]} ({sc})"
    for name, a, b in zip(("params", "meta-gradients", "mu", "nu"), flat[:4], leaf[:4]):
        assert len(a) == len(b), name
        for i, (x, y) in enumerate(zip(a, b)):
            if name == "meta-gradients" and (x is None or y is None):
                # a leaf whose gradient is structurally absent in one layout may be an exact zero in the other
                z = x if x is not None else y
                assert z is None or not z.any(), f"{name}[{i}] ({sc})"
                continue
            assert x.shape == y.shape and x.dtype == y.dtype and torch.equal(x, y), f"{name}[{i}] differs ({sc})"
