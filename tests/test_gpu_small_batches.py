"""Explain-path batch sizes (a handful of candidates, a few rows per step): the kernels that serve them --
the per-row strip merge (cx_merge_strips), the tcgen05 fused pass from 1 row on (umma_min_rows), the skinny GEMM of
ConvE's Linear forward (skinny_fc) -- against the kernels they replaced (A/B through kp_set_option) and the oracle."""
import numpy as np
import pytest
import torch

from oracle import kelpie_oracle as ko
from tests.golden_util import load, seed_all

pytestmark = pytest.mark.gpu
RTOL = 1e-4  # mimic rows: max |diff| <= RTOL * max |row| (north_star's fp32 tolerance)


def _ctx(kind, z, w):
    from kelpie_b200 import runtime
    return runtime.Context(kind, z["w_ent"], z["w_rel"], norm=2, conve=dict(w.conve) if kind == "ConvE" else None)


def _jobs(rng, N, R, D, sizes, scale):
    out = []
    for T in sizes:
        facts = []
        for _ in range(T):
            x, r = int(rng.integers(0, N)), int(rng.integers(0, R))
            facts.append((N, r, x) if rng.random() < 0.6 else (x, r, N))
        out.append((facts, (rng.random(D) * scale).astype(np.float32)))
    return out


def _run(ctx, kind, kg, hp, jobs, **opts):
    from kelpie_b200 import plans, runtime
    for k, v in opts.items():
        ctx.set_option(k, v)
    seed_all(5)
    b = plans.Batch(kind, kg.num_entities, kg.num_relations, hp)
    for f, i in jobs:
        b.add(f, i)
    return ctx.post_train(runtime.make_hp(kind, hp), **b.arrays()).cpu().numpy()


def _oracle(kind, w, kg, hp, jobs):
    seed_all(5)
    return np.stack([ko.post_train(w, kg, torch.from_numpy(i).view(1, -1), f, hp)[-1].numpy() for f, i in jobs])


def _rel(a, b):
    return (np.abs(a - b) / np.maximum(np.abs(b).max(axis=1, keepdims=True), 1e-30)).max()


@pytest.mark.parametrize("sizes", [[1], [3, 5, 2, 4], [9, 1, 1, 7, 2, 6]])
def test_complex_small_batches_match_oracle_and_previous_kernels(sizes):
    z, meta, kg, w, order = load("ComplEx")
    ctx = _ctx("ComplEx", z, w)
    hp = dict(meta["hp"], epochs=6)
    jobs = _jobs(np.random.default_rng(len(sizes)), kg.num_entities, kg.num_relations, w.dim, sizes, 1e-3)
    want = _oracle("ComplEx", w, kg, hp, jobs)
    new = _run(ctx, "ComplEx", kg, hp, jobs, umma_min_rows=1, cx_merge=1)
    assert _rel(new, want) <= RTOL
    # the per-row merge folds the strips in the order and with the weights of the serial merge: identical rows
    serial = _run(ctx, "ComplEx", kg, hp, jobs, umma_min_rows=1, cx_merge=0)
    np.testing.assert_array_equal(new, serial)
    # CUDA-core pass for < 32 rows (the earlier dispatch), merged either way
    simt = _run(ctx, "ComplEx", kg, hp, jobs, umma_min_rows=32, cx_merge=0)
    np.testing.assert_array_equal(simt, _run(ctx, "ComplEx", kg, hp, jobs, umma_min_rows=32, cx_merge=1))
    assert _rel(simt, want) <= RTOL and _rel(new, simt) <= RTOL
    ctx.set_option("umma_min_rows", 1)


@pytest.mark.parametrize("sizes", [[2], [3, 6, 1, 4], [12, 30, 25, 40]])
def test_conve_skinny_linear_forward_matches_tiled_gemm_and_oracle(sizes):
    """The last case has 64..127 (s, p) pairs in a step: still below the tcgen05 Linear layer's 128 rows."""
    z, meta, kg, w, order = load("ConvE")
    ctx = _ctx("ConvE", z, w)
    hp = dict(meta["hp"], epochs=4)
    jobs = _jobs(np.random.default_rng(7 + len(sizes)), kg.num_entities, kg.num_relations, w.dim, sizes, 1.0)
    want = _oracle("ConvE", w, kg, hp, jobs)
    skinny = _run(ctx, "ConvE", kg, hp, jobs, skinny_fc=1)
    tiled = _run(ctx, "ConvE", kg, hp, jobs, skinny_fc=0)
    assert _rel(skinny, want) <= RTOL and _rel(tiled, want) <= RTOL
    assert _rel(skinny, tiled) <= 1e-5
    np.testing.assert_array_equal(skinny, _run(ctx, "ConvE", kg, hp, jobs, skinny_fc=1))  # fixed-order reduction
    # inference through the same Linear layer: scores of a few queries
    q = np.array([[int(f[0][2] if f[0][0] == kg.num_entities else f[0][0]), int(f[0][1]), 0] for f, _ in jobs], dtype=np.int64)
    ctx.set_option("skinny_fc", 1)
    a = ctx.all_scores(q).cpu().numpy()
    ctx.set_option("skinny_fc", 0)
    b = ctx.all_scores(q).cpu().numpy()
    ctx.set_option("skinny_fc", 1)
    assert np.abs(a - b).max() <= 1e-5 * max(np.abs(b).max(), 1e-30)


def test_replay_self_check_leaves_every_generator_untouched():
    """plans.HostReplay.available() runs lazily in the middle of a run (first plan drawn): the CPU, CUDA and numpy
    generators must come out of it exactly as they went in (the reference's mimic rows are drawn on the CUDA one)."""
    from kelpie_b200 import plans
    torch.manual_seed(42); np.random.seed(42)
    torch.rand(3); torch.rand(3, device="cuda"); np.random.random(3)
    before = (torch.get_rng_state().clone(), torch.cuda.get_rng_state().clone(), np.random.get_state())
    saved, plans.HostReplay._ok = plans.HostReplay._ok, None
    try:
        assert plans.HostReplay.available()
    finally:
        plans.HostReplay._ok = saved if saved is not None else plans.HostReplay._ok
    assert torch.equal(before[0], torch.get_rng_state()) and torch.equal(before[1], torch.cuda.get_rng_state())
    after = np.random.get_state()
    assert np.array_equal(before[2][1], after[1]) and before[2][2:] == after[2:]
