"""CPU tests of the host-side logic: shard/partition math, the pytree shim, step-counter mirroring, and the
N > 1 paths under gloo with world_size 2 (flat shards tile the state exactly; the task-parallel meta-gradient
all-reduce equals the single-process mean)."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import assert_bits_equal


def test_shard_range_tiles_and_aligns():
    from torchopt_b200.parallel import shard_range
    for n in (0, 1, 7, 8, 9, 1000, 2 ** 20 + 3, 2 ** 30):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (lo, hi), (lo2, _) in zip(edges, edges[1:]):
                assert hi == lo2 and lo <= hi
            for lo, hi in edges[:-1]:
                assert hi % 8 == 0 or hi == n
            sizes = [hi - lo for lo, hi in edges]
            if n >= 8 * world:
                assert max(sizes) - min(sizes) < 16
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_partition_leaves_and_task_slice():
    from torchopt_b200.parallel import partition_leaves, task_slice
    sizes = [9408, 64, 64, 36864, 64, 64, 2359296, 512, 512, 512000, 1000]
    for world in (1, 2, 4, 8):
        bins = partition_leaves(sizes, world)
        assert sorted(i for b in bins for i in b) == list(range(len(sizes)))
        loads = [sum(sizes[i] for i in b) for b in bins]
        assert max(loads) <= max(max(sizes), sum(sizes) / world * 1.34 + max(sizes) * (world == 1))
    assert partition_leaves(sizes, 3) == partition_leaves(sizes, 3)  # deterministic
    for num, world in ((32, 8), (32, 3), (5, 8), (0, 2)):
        got = [list(task_slice(num, r, world)) for r in range(world)]
        assert sum(got, []) == list(range(num))
        assert max(map(len, got)) - min(map(len, got)) <= 1


def test_pytree_shim_none_is_leaf():
    from collections import namedtuple
    from torchopt_b200 import _pytree
    NT = namedtuple("NT", ["a", "b"])
    tree = {"z": [1, None, (2, 3)], "a": NT(4, {"k": 5}), "m": ()}
    leaves, spec = _pytree.tree_flatten(tree)
    assert leaves == [4, 5, 1, None, 2, 3]          # dict keys sorted; None is a leaf
    assert spec.num_leaves == 6
    back = _pytree.tree_unflatten(spec, leaves)
    assert back == tree and isinstance(back["a"], NT)
    doubled = _pytree.tree_map(lambda x, y: None if x is None else x + y, tree, tree)
    assert doubled["z"] == [2, None, (4, 6)] and doubled["a"].b["k"] == 10
    with pytest.raises(ValueError):
        _pytree.tree_map(lambda x, y: x, tree, {"z": 1})


def test_step_counters_host_mirror_cpu():
    """inc_count semantics of transform/utils.py:111-122 with the host mirror: +1 where a gradient exists,
    saturating at INT64_MAX, untouched where the gradient is None; counters stay 0-dim int64 tensors."""
    from torchopt_b200.accelerated_op.adam_op import INT64_MAX, count_as_int
    from torchopt_b200.transform import inc_count_flat, make_counts
    dev = torch.device("cpu")
    counts = make_counts([0, 5, None, INT64_MAX], [dev, dev, None, dev])
    assert counts[2] is None and all(c.dim() == 0 and c.dtype == torch.int64 for c in counts if c is not None)
    g = [torch.ones(1), None, None, torch.ones(1)]
    new = inc_count_flat(g, counts)
    assert [None if c is None else int(c) for c in new] == [1, 5, None, INT64_MAX]
    assert new[1] is counts[1]
    assert count_as_int(new[0]) == 1 and new[0]._b200_host_count == 1
    bare = [torch.tensor(41), torch.tensor(7)]       # user-built counters without a mirror
    new = inc_count_flat([torch.ones(1), torch.ones(1)], bare)
    assert [int(c) for c in new] == [42, 8] and count_as_int(3) == 3
    # an in-place edit of a tagged counter bumps its version: the stale mirror must not be used
    new[0].fill_(100)
    assert count_as_int(new[0]) == 100 and new[0]._b200_host_count == 100
    again = inc_count_flat([torch.ones(1), torch.ones(1)], new)
    assert [int(c) for c in again] == [101, 9]


def test_transform_validation_and_none_leaves_cpu():
    import torchopt_b200 as tb
    with pytest.raises(ValueError):
        tb.scale_by_accelerated_adam(b1=1.0)
    with pytest.raises(ValueError):
        tb.scale_by_accelerated_adam(eps=-1.0)
    with pytest.raises(NotImplementedError):
        tb.adam(use_accelerated_op=False)
    # empty / None-only pytrees never reach a kernel (reference tests/test_alias.py:35-79)
    tx = tb.scale_by_accelerated_adam()
    for params in ((), [], {}, {"a": None}, (None, None)):
        state = tx.init(params)
        updates, state = tx.update(params, state, inplace=False)
        assert updates == params
    opt = tb.adam(1e-3)
    st = opt.init([])
    assert opt.update([], st, params=[], inplace=True)[0] == []


# ---------------------------------------------------------------------------------------------------------------
# world_size-2 gloo
# ---------------------------------------------------------------------------------------------------------------
def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, tmp: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import adam_oracle as ao
        from torchopt_b200.parallel import allreduce_meta_grads, shard_range, task_slice
        # (1) flat-state sharding: every rank processes its contiguous range with the same (oracle) arithmetic;
        #     the concatenation must equal the unsharded result bit-for-bit, with no communication in the op.
        n = 100_003
        rng = np.random.default_rng(7)
        g = (rng.standard_normal(n) * 1e-2).astype(np.float32)
        mu = (rng.standard_normal(n) * 1e-2).astype(np.float32)
        nu = (rng.random(n) * 1e-4).astype(np.float32)
        lo, hi = shard_range(n, rank, world)
        orc = ao.COracle()
        _, _, u = ao.fused_forward(orc, g[lo:hi], mu[lo:hi], nu[lo:hi], 0.9, 0.999, 1e-8, 0.0, 2)
        np.save(os.path.join(tmp, f"u_{rank}.npy"), u)
        # (2) task-parallel meta-gradient: rank-local SUM of per-task grads, one all-reduce of one bucket,
        #     pre-scaled by 1/num_tasks == mean over the whole meta-batch.
        num_tasks = 7
        torch.manual_seed(0)
        per_task = [[torch.randn(5, 3), torch.randn(11), torch.randn(4, dtype=torch.float64)]
                    for _ in range(num_tasks)]                                        # same on every rank
        local = [sum(per_task[t][k] for t in task_slice(num_tasks, rank, world)) for k in range(3)]
        local.insert(2, None)                                                         # an unused parameter
        out = allreduce_meta_grads(local, scale=1.0 / num_tasks)
        want = [torch.stack([per_task[t][k] for t in range(num_tasks)]).mean(0) for k in range(3)]
        assert out[2] is None
        for a, b in zip(out[:2], want[:2]):
            torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-6)
        # an fp64 meta-gradient is reduced in fp64 (its own bucket), not truncated to fp32
        assert out[3].dtype == torch.float64
        torch.testing.assert_close(out[3], want[2], rtol=1e-14, atol=1e-14)
        # (3) the same through GradBucket: .grad views of one flat buffer, backward accumulates in place,
        #     the all-reduce runs on the buffer itself
        from torchopt_b200.parallel import GradBucket
        torch.manual_seed(1)
        w = [torch.randn(4, 3, requires_grad=True), torch.randn(6, requires_grad=True)]
        xs = [torch.randn(3) for _ in range(num_tasks)]
        bucket = GradBucket(w)
        bucket.zero()
        for t in task_slice(num_tasks, rank, world):
            (((w[0] @ xs[t]).sum() * (t + 1) + (w[1] ** 2).sum() * t) / num_tasks).backward()
        assert w[0].grad.data_ptr() == bucket.views[0].data_ptr()     # accumulated in place, not replaced
        bucket.allreduce()
        ref = [torch.zeros_like(x) for x in w]
        for t in range(num_tasks):
            gs = torch.autograd.grad(((w[0] @ xs[t]).sum() * (t + 1) + (w[1] ** 2).sum() * t) / num_tasks, w)
            ref = [a + b for a, b in zip(ref, gs)]
        for a, b in zip(w, ref):
            torch.testing.assert_close(a.grad, b, rtol=1e-5, atol=1e-6)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    from oracle import adam_oracle as ao
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    n = 100_003
    rng = np.random.default_rng(7)
    g = (rng.standard_normal(n) * 1e-2).astype(np.float32)
    mu = (rng.standard_normal(n) * 1e-2).astype(np.float32)
    nu = (rng.random(n) * 1e-4).astype(np.float32)
    _, _, want = ao.fused_forward(ao.COracle(), g, mu, nu, 0.9, 0.999, 1e-8, 0.0, 2)
    got = np.concatenate([np.load(tmp_path / f"u_{r}.npy") for r in range(world)])
    assert_bits_equal(got, want, "sharded == unsharded")


def test_allreduce_meta_grads_single_process():
    from torchopt_b200.parallel import allreduce_meta_grads
    g = [torch.ones(3), None, torch.full((2, 2), 4.0)]
    out = allreduce_meta_grads(g, scale=0.5)
    assert out[1] is None and torch.equal(out[0], torch.full((3,), 0.5)) and torch.equal(out[2], torch.full((2, 2), 2.0))
    assert allreduce_meta_grads([None]) == [None]


def test_zeros_like_flat_matches_zeros_like():
    """Flat-carved moments have exactly the metadata torch.zeros_like gives (sizes, strides of dense leaves, dtype,
    leaf-ness, requires_grad) and 128-byte aligned offsets inside their buffer."""
    from torchopt_b200.transform import zeros_like_flat
    leaves = [torch.randn(5, 3), None, torch.randn(4, 6).t(), torch.randn(7, dtype=torch.float64), torch.randn(2, 3, 4),
              torch.randn(2, 2, dtype=torch.float64), torch.randn(2, 3, 4, 5).to(memory_format=torch.channels_last)]
    for rg in (False, True):
        out = zeros_like_flat(leaves, requires_grad=rg)
        base = {}
        for a, b in zip(leaves, out):
            if a is None:
                assert b is None
                continue
            r = torch.zeros_like(a)
            assert b.shape == r.shape and b.stride() == r.stride() and b.dtype == r.dtype
            assert b.is_leaf and b.requires_grad == rg and bool((b == 0).all())
            first = base.setdefault(b.dtype, b.data_ptr())
            assert (b.data_ptr() - first) % 128 == 0
        if rg:   # independent leaves: a gradient flows into one without touching the others
            (out[0] * 2).sum().backward()
            assert out[0].grad is not None and out[2].grad is None


def test_grad_bucket_single_process():
    from torchopt_b200.parallel import GradBucket
    w = [torch.randn(3, 2, requires_grad=True), torch.randn(5, requires_grad=True)]
    bucket = GradBucket(w)
    (w[0].sum() * 2 + (w[1] * 3).sum()).backward()
    (w[0].sum() * 2 + (w[1] * 3).sum()).backward()
    assert torch.equal(w[0].grad, torch.full((3, 2), 4.0)) and torch.equal(w[1].grad, torch.full((5,), 6.0))
    flat = next(iter(bucket.buffers.values()))
    assert flat.sum() == 4.0 * 6 + 6.0 * 5              # the views ARE the buffer
    bucket.allreduce(scale=0.5)
    assert torch.equal(w[0].grad, torch.full((3, 2), 2.0))
    for p in w:
        p.grad = None
    bucket.zero()
    assert w[0].grad is bucket.views[0] and float(flat.abs().sum()) == 0.0


def test_pytree_helpers_leave_no_reference_cycles():
    """tree_flatten / tree_unflatten / tree_map must not create garbage: a self-referential closure would keep the leaves
    -- tensors with their autograd graphs -- alive until the next gc pass (memory held across unrolled steps, and stale
    gradient-accumulator nodes that break CUDA-graph capture after an eager run)."""
    import gc
    import weakref

    import torch
    from torchopt_b200 import _pytree
    gc.collect()
    gc.disable()
    try:
        a = torch.randn(5, requires_grad=True)
        t = a * 2
        ref = weakref.ref(t)
        tree = {"x": [t, None], "y": (t + 1,)}
        leaves, spec = _pytree.tree_flatten(tree)
        back = _pytree.tree_unflatten(spec, leaves)
        mapped = _pytree.tree_map(lambda v: v, tree)
        del t, tree, leaves, back, mapped
        assert ref() is None, "a pytree helper kept the leaves alive (reference cycle)"
    finally:
        gc.enable()
    assert gc.collect() == 0


# ---------------------------------------------------------------------------------------------------------------
# round 2: flat Adam state plumbing, bench helpers, drop-in staging -- everything that needs no GPU
# ---------------------------------------------------------------------------------------------------------------
def test_flat_adam_state_layout_and_roundtrip_through_the_reference_layout():
    """FlatAdamState <-> ScaleByAdamState: 128-byte aligned leaf offsets per (device, dtype) group, per-leaf views that
    alias the flat buffers, host counts carried as tagged count tensors, and a differentiable re-pack."""
    import torchopt_b200 as tb
    from torchopt_b200.optim import FlatAdamState
    leaves = [torch.zeros(5), torch.zeros(3, 4), torch.zeros((), dtype=torch.float64), torch.zeros(70), torch.zeros(0)]
    st = FlatAdamState(leaves, moment_requires_grad=True)
    by_dtype = {g.dtype: g for g in st.groups}
    assert by_dtype[torch.float32].offsets == [0, 32, 64, 160] and by_dtype[torch.float32].total == 160
    assert by_dtype[torch.float64].offsets == [0] and by_dtype[torch.float64].total == 16
    assert st.matches(leaves) and not st.matches(leaves[:-1]) and not st.matches([t.double() for t in leaves])
    st.counts = [3, 3, 0, 3, 1]
    with torch.no_grad():
        by_dtype[torch.float32].mu.copy_(torch.arange(160.0))
    leaf = st.to_leaf_state()
    assert isinstance(leaf, tb.ScaleByAdamState)
    assert [tuple(m.shape) for m in leaf.mu] == [(5,), (3, 4), (), (70,), (0,)]
    assert torch.equal(leaf.mu[1], torch.arange(32.0, 44.0).view(3, 4)) and leaf.mu[3][0] == 64
    assert leaf.mu[1].data_ptr() == by_dtype[torch.float32].mu.data_ptr() + 4 * 32          # views, not copies
    assert [int(c) for c in leaf.count] == [3, 3, 0, 3, 1]
    assert all(getattr(c, "_b200_host_count") == v for c, v in zip(leaf.count, st.counts))  # no sync needed to read them
    back = FlatAdamState.from_leaf_state(leaves, leaf)
    assert back.counts == st.counts
    g32 = next(g for g in back.groups if g.dtype == torch.float32)
    assert torch.equal(g32.mu.detach()[:5], torch.arange(5.0)) and torch.equal(g32.mu.detach()[64:134], torch.arange(64.0, 134.0))
    assert not g32.mu.detach()[5:32].any()                                                   # padding re-packed as zeros
    # the re-pack is differentiable w.r.t. the tensors the caller holds (gradients of an extracted state keep flowing)
    (gsum,) = torch.autograd.grad(g32.mu[32:44].sum() * 2.0, [leaf.mu[1]])
    assert torch.equal(gsum, torch.full((3, 4), 2.0))
    with pytest.raises(ValueError):
        FlatAdamState.from_leaf_state([t.clone() for t in leaves[:2]] + [torch.zeros(2)] + leaves[3:],
                                      tb.ScaleByAdamState(mu=leaf.mu, nu=leaf.nu, count=leaf.count))


def test_optimizer_state_property_switches_between_layouts():
    import torchopt_b200 as tb
    from torchopt_b200.optim import FlatAdamState
    opt = tb.FuncAdam(1e-3, weight_decay=1e-2)
    assert opt.state is None and opt.state_dict() is None
    leaves = [torch.zeros(4), torch.zeros(2, 2)]
    flat = opt._flat_state(leaves, False)
    assert isinstance(flat, FlatAdamState) and opt._leaf is None
    state = opt.state                               # materialises the reference layout: chain tuple with one Adam state
    adam = [s for s in state if isinstance(s, tb.ScaleByAdamState)]
    assert len(adam) == 1 and len(state) == 3 and len(adam[0].mu) == 2
    assert opt._flat_state(leaves, False) is not flat, "a state somebody read is re-packed from exactly those tensors"
    assert opt._leaf is None
    opt.load_state_dict(state)
    assert opt._flat is None and opt.state is state
    opt.state = None
    assert opt._flat is None and opt._leaf is None
    with pytest.raises(ValueError):
        opt._flat_state(leaves, False)
        opt._flat_state([torch.zeros(5), torch.zeros(2, 2)], False)


def test_bench_helpers():
    import bench
    assert bench.e2e_sample_elems(1 << 20, 1 << 30, 1) == 1 << 20
    assert bench.e2e_sample_elems(0, 1 << 10, 1) == 1 << 10
    m = bench.e2e_sample_elems(0, 1 << 40, 8)
    assert (1 << 24) <= m < (1 << 40) and m & (m - 1) == 0
    import subprocess as sp
    topo = ("\tGPU0\tGPU1\tNIC0\tCPU Affinity\tNUMA Affinity\tGPU NUMA ID\n"
            "GPU0\t X \tNV18\tSYS\t0-3,8-9\t0\t\tN/A\nGPU1\tNV18\t X \tSYS\t4-7\t1\t\tN/A\n\nLegend:\n")

    class Res:
        stdout = topo

    real = sp.run
    sp.run = lambda *a, **k: Res()
    try:
        assert bench.local_cpu_affinity(0) == [0, 1, 2, 3, 8, 9] and bench.local_cpu_affinity(1) == [4, 5, 6, 7]
        assert bench.local_cpu_affinity(2) is None
    finally:
        sp.run = real


def test_optree_stand_in_follows_optree_conventions():
    """tests/dropin/shims/optree (test-only stand-in used by the drop-in harness): sorted dict keys, None is an internal
    node unless none_is_leaf, namedtuples keep their type, tree_transpose, flatten_up_to."""
    import importlib.util
    import sys
    from collections import OrderedDict, namedtuple
    shim = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "dropin", "shims")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "optree" or k.startswith("optree.")}
    sys.path.insert(0, shim)
    try:
        import optree
        Pt = namedtuple("Pt", "x y")
        tree = {"b": (1, None, [2, 3]), "a": Pt(4, {"z": 5})}
        leaves, spec = optree.tree_flatten(tree)
        assert leaves == [4, 5, 1, 2, 3] and spec.num_leaves == 5
        assert optree.tree_unflatten(spec, leaves) == tree
        leaves_nil, spec_nil = optree.tree_flatten(tree, none_is_leaf=True)
        assert leaves_nil == [4, 5, 1, None, 2, 3] and spec_nil != spec
        assert optree.tree_map(lambda a, b: a + b, (1, [2]), (10, [20])) == (11, [22])
        assert optree.tree_map(lambda a: None if a is None else a * 2, (1, None), none_is_leaf=True) == (2, None)
        assert list(optree.tree_flatten(OrderedDict([("b", 1), ("a", 2)]))[0]) == [1, 2]       # insertion order kept
        outer, inner = optree.tree_structure([0, 0]), optree.tree_structure((0, 0, 0))
        assert optree.tree_transpose(outer, inner, [(1, 2, 3), (4, 5, 6)]) == ([1, 4], [2, 5], [3, 6])
        assert spec.flatten_up_to({"b": ("B0", None, ["l0", "l1"]), "a": Pt("px", {"z": "pz"})}) == ["px", "pz", "B0", "l0", "l1"]
        with pytest.raises(ValueError):
            spec.flatten_up_to({"b": (1, None, [2]), "a": Pt(4, {"z": 5})})
        from optree.typing import PyTree, PyTreeTypeVar   # noqa: F401  what torchopt/typing.py:36 imports
        assert PyTree[int] is not None
    finally:
        sys.path.remove(shim)
        for k in [k for k in sys.modules if k == "optree" or k.startswith("optree.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_dropin_archive_holds_the_reference_package_and_our_edits():
    """oracle/_ref/dropin.tar.gz (built by __graft_entry__.build() when /root/reference is mounted)."""
    import tarfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    archive = os.path.join(root, "oracle", "_ref", "dropin.tar.gz")
    if not os.path.isfile(archive):
        pytest.skip("no staged archive (reference not mounted and no earlier build)")
    with tarfile.open(archive) as tar:
        names = set(tar.getnames())
        for need in ("plain/torchopt/accelerated_op/adam_op.py", "fused/torchopt/transform/scale_by_adam.py",
                     "reftests/test_alias.py", "reftests/test_optim.py", "reftests/helpers.py", "reftests/conftest.py"):
            assert need in names, need
        assert not any(n.endswith(".so") for n in names), "native modules are copied in at test time, never archived"
        fused_op = tar.extractfile("fused/torchopt/accelerated_op/adam_op.py").read().decode()
        plain_op = tar.extractfile("plain/torchopt/accelerated_op/adam_op.py").read().decode()
        fused_tr = tar.extractfile("fused/torchopt/transform/scale_by_adam.py").read().decode()
        plain_tr = tar.extractfile("plain/torchopt/transform/scale_by_adam.py").read().decode()
    assert "from torchopt._C import adam_op" in fused_op and "class FusedOp" in fused_op and "def multi" in fused_op
    assert "FusedOp" not in plain_op and "op.multi(" in fused_tr and "op.multi(" not in plain_tr
    if os.path.isfile("/root/reference/torchopt/accelerated_op/adam_op.py"):
        assert plain_op == open("/root/reference/torchopt/accelerated_op/adam_op.py").read()
