"""Edge cases of the post-training kernels against the oracle on identical draws: multi-step
epochs (rows > batch_size), mixed static / per-epoch batches, L1 TransE, empty candidates,
self-loop facts, ragged batches; plus size-independent properties of the rank kernel."""
import numpy as np
import pytest
import torch

from oracle import kelpie_oracle as ko
from tests.golden_util import load, seed_all

pytestmark = pytest.mark.gpu
RTOL = 1e-4  # mimic rows: max |diff| <= RTOL * max |row|


def _ctx(kind, z, w, norm=2):
    from kelpie_b200 import runtime
    conve = dict(w.conve) if kind == "ConvE" else None
    return runtime.Context(kind, z["w_ent"], z["w_rel"], norm=norm, conve=conve)


def _both(kind, z, w, kg, hp, jobs, ctx):
    """Oracle rows and CUDA rows for the same jobs, same generator state."""
    from kelpie_b200 import plans, runtime
    N, R = kg.num_entities, kg.num_relations
    seed_all(11)
    want = []
    for facts, init in jobs:
        table = ko.post_train(w, kg, torch.from_numpy(init).view(1, -1), facts, hp)
        want.append(table[-1].numpy())
    seed_all(11)
    b = plans.Batch(kind, N, R, hp)
    for facts, init in jobs:
        b.add(facts, init)
    got = ctx.post_train(runtime.make_hp(kind, hp), **b.arrays()).cpu().numpy()
    if kind == "TransE":  # the compact index tables (ABI 2) drive the same arithmetic: bit-identical rows
        compact = ctx.post_train(runtime.make_hp(kind, hp), **b.arrays(compact=True)).cpu().numpy()
        assert np.array_equal(compact, got)
    return got, np.stack(want)


def _jobs(rng, N, R, D, sizes, scale):
    out = []
    for T in sizes:
        facts = []
        for _ in range(T):
            x, r = int(rng.integers(0, N)), int(rng.integers(0, R))
            facts.append((N, r, x) if rng.random() < 0.5 else (x, r, N))
        out.append((facts, (rng.random(D) * scale).astype(np.float32)))
    return out


def _assert_rows(got, want):
    scale = np.maximum(np.abs(want).max(axis=1, keepdims=True), 1e-30)
    assert (np.abs(got - want) / scale).max() <= RTOL


@pytest.mark.parametrize("norm", [1, 2])
def test_transe_multi_step_ragged_and_empty(norm):
    z, meta, kg, w, order = load("TransE")
    w.norm = norm
    hp = dict(meta["hp"], batch_size=4, epochs=7)  # 2T > batch_size -> several dependent steps per epoch
    rng = np.random.default_rng(3)
    jobs = _jobs(rng, kg.num_entities, kg.num_relations, w.dim, [0, 1, 2, 5, 9, 3], 0.3)
    jobs[3][0][0] = (kg.num_entities, 1, kg.num_entities)  # self-loop fact (M, r, M)
    got, want = _both("TransE", z, w, kg, hp, jobs, _ctx("TransE", z, w, norm))
    np.testing.assert_array_equal(got[0], jobs[0][1])  # no facts: init row untouched
    _assert_rows(got, want)


def test_complex_multi_step_and_mixed_batches():
    z, meta, kg, w, order = load("ComplEx")
    hp = dict(meta["hp"], batch_size=6, epochs=5)  # jobs with 2T <= 6 are static, the others permuted per epoch
    rng = np.random.default_rng(4)
    jobs = _jobs(rng, kg.num_entities, kg.num_relations, w.dim, [0, 2, 3, 7, 10, 1], 1e-3)
    jobs[4][0][1] = (kg.num_entities, 2, kg.num_entities)
    got, want = _both("ComplEx", z, w, kg, hp, jobs, _ctx("ComplEx", z, w))
    np.testing.assert_array_equal(got[0], jobs[0][1])
    _assert_rows(got, want)


@pytest.mark.parametrize("opt", ["Adam", "SGD"])
def test_complex_other_optimizers_and_n3(opt):
    z, meta, kg, w, order = load("ComplEx")
    hp = dict(meta["hp"], optimizer_name=opt, lr=0.01, epochs=4, regularizer_weight=0.05)
    rng = np.random.default_rng(5)
    jobs = _jobs(rng, kg.num_entities, kg.num_relations, w.dim, [3, 6], 0.2)
    got, want = _both("ComplEx", z, w, kg, hp, jobs, _ctx("ComplEx", z, w))
    _assert_rows(got, want)


def test_conve_multi_step_pairs():
    z, meta, kg, w, order = load("ConvE")
    hp = dict(meta["hp"], batch_size=3, epochs=4)  # pairs > batch_size -> several steps per epoch
    rng = np.random.default_rng(6)
    jobs = _jobs(rng, kg.num_entities, kg.num_relations, w.dim, [0, 4, 6, 2], 1.0)
    got, want = _both("ConvE", z, w, kg, hp, jobs, _ctx("ConvE", z, w))
    np.testing.assert_array_equal(got[0], jobs[0][1])
    _assert_rows(got, want)


@pytest.mark.parametrize("kind", ["TransE", "ComplEx", "ConvE"])
def test_rank_kernel_properties(kind):
    """Size-independent properties: the fused rank equals a count over the materialised scores,
    filtering never worsens a rank, and filtering everything leaves rank 1."""
    from kelpie_b200 import runtime
    z, meta, kg, w, order = load(kind)
    ctx = _ctx(kind, z, w)
    N = kg.num_entities
    rng = np.random.default_rng(8)
    Q = 150
    triples = np.stack([rng.integers(0, N, Q), rng.integers(0, 2 * kg.num_relations, Q), rng.integers(0, N, Q)], 1)
    sc = ctx.all_scores(triples)
    t = sc.gather(1, torch.as_tensor(triples[:, 2], device=sc.device).view(-1, 1))
    better = (sc < t) if kind == "TransE" else (sc > t)
    ties = (sc == t)
    empty = np.zeros(Q + 1, dtype=np.int64)
    ts, bs, rk, cnt = ctx.filtered_rank(triples, runtime.RANK_MODEL, flt_off=empty, counters=True)
    assert torch.equal(cnt[:, 0].long(), better.sum(1))
    assert torch.equal(rk, better.sum(1) + ties.sum(1))  # ties (incl. the target itself) count against
    torch.testing.assert_close(ts, t.view(-1), rtol=1e-5, atol=1e-6)
    # random filters (ragged, some empty): rank can only improve
    lens = rng.integers(0, 40, Q)
    off = np.zeros(Q + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    ids = np.concatenate([np.sort(rng.choice(N, n, replace=False)) for n in lens]).astype(np.int32)
    _, _, rk_f = ctx.filtered_rank(triples, runtime.RANK_MODEL, flt_off=off, flt_ids=ids)
    assert bool((rk_f <= rk).all())
    # everything filtered: only the (restored) target remains
    allids = np.tile(np.arange(N, dtype=np.int32), Q)
    off_all = np.arange(Q + 1, dtype=np.int64) * N
    _, _, rk_all = ctx.filtered_rank(triples, runtime.RANK_MODEL, flt_off=off_all, flt_ids=allids)
    assert bool((rk_all == 1).all())


def test_post_training_is_deterministic_and_batch_order_independent():
    from kelpie_b200 import plans, runtime
    z, meta, kg, w, order = load("ComplEx")
    ctx = _ctx("ComplEx", z, w)
    hp = dict(meta["hp"], epochs=5)
    rng = np.random.default_rng(9)
    jobs = _jobs(rng, kg.num_entities, kg.num_relations, w.dim, [3, 5, 8, 2, 6, 4, 7, 9] * 6, 1e-3)

    def run(js):
        b = plans.Batch("ComplEx", kg.num_entities, kg.num_relations, hp)
        for f, i in js:
            b.add(f, i)
        return ctx.post_train(runtime.make_hp("ComplEx", hp), **b.arrays()).cpu().numpy()

    a = run(jobs)
    np.testing.assert_array_equal(a, run(jobs))          # idempotent / deterministic
    perm = rng.permutation(len(jobs))
    b = run([jobs[i] for i in perm])
    scale = np.abs(a).max(axis=1, keepdims=True)
    assert (np.abs(b - a[perm]) / scale[perm]).max() < 1e-5  # a candidate does not depend on its neighbours


@pytest.mark.parametrize("kind", ["TransE", "ComplEx", "ConvE"])
@pytest.mark.parametrize("Q", [1, 3, 8])
def test_streaming_pass_equals_tile_pass(kind, Q):
    """Few-query HBM-streaming kernel (kp_stream.cu) vs the 64-query tile kernel (kp_pass.cu)."""
    from kelpie_b200 import runtime
    z, meta, kg, w, order = load(kind)
    ctx = _ctx(kind, z, w)
    N = kg.num_entities
    rng = np.random.default_rng(Q)
    triples = np.stack([rng.integers(0, N, Q), rng.integers(0, 2 * kg.num_relations, Q), rng.integers(0, N, Q)], 1)
    mimic = (rng.standard_normal((Q, w.dim)) * 0.1).astype(np.float32)
    triples[0, 0] = N  # one query whose lhs is its mimic row
    lens = rng.integers(0, 30, Q)
    off = np.zeros(Q + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    ids = np.concatenate([np.sort(rng.choice(N + 1, n, replace=False)) for n in lens] + [np.zeros(0, np.int64)]).astype(np.int32)
    res = {}
    for tile in (0, 1):
        ctx.set_option("force_tile", tile)
        sc = ctx.all_scores(triples, mimic_rows=mimic).cpu().numpy()
        out = [t.cpu().numpy() for t in ctx.filtered_rank(triples, runtime.RANK_ENGINE_MIN if kind == "TransE" else runtime.RANK_ENGINE_MAX,
                                                            mimic_rows=mimic, flt_off=off, flt_ids=ids if len(ids) else None, counters=True)]
        res[tile] = (sc, out)
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=2e-5, atol=1e-6)
    np.testing.assert_array_equal(res[0][1][2], res[1][1][2])   # ranks
    np.testing.assert_array_equal(res[0][1][3], res[1][1][3])   # counters
    np.testing.assert_allclose(res[0][1][0], res[1][1][0], rtol=2e-5, atol=1e-6)


@pytest.mark.gpu
def test_device_built_filter_csr_equals_dict_upload():
    """kp_filter_build (sort / unique / segment on the device) leaves the same resident CSR as
    kp_filter_upload fed by the dict walk -- on the real DBpedia50 filter (train + valid + test, both directions),
    after dataset edits, and for the empty case."""
    import os
    from kelpie_b200 import runtime
    from kelpie_b200.data import Dataset
    from tests.golden_util import GOLDEN
    ds = Dataset.from_npz(os.path.join(GOLDEN, "dbpedia50_ids.npz"), name="DBpedia50")
    ds.remove_training_triples([tuple(int(x) for x in t) for t in ds.training_triples[:50]])
    ds.add_training_triples([(5, 3, 9), (5, 3, 9), (17, 0, 2)])
    N, R2 = ds.num_entities, 2 * ds.num_relations
    ent, rel = torch.zeros(N, 8), torch.zeros(R2, 8)
    a, b = runtime.Context("TransE", ent, rel), runtime.Context("TransE", ent, rel)
    a.upload_filter(ds.to_filter)
    b.build_filter(ds.filter_facts())
    for x, y in zip(a.download_filter(), b.download_filter()):
        assert x.dtype == y.dtype and np.array_equal(x, y)
    keys, off, ids = b.download_filter()
    assert len(keys) > 30000 and (np.diff(keys) > 0).all() and off[-1] == len(ids)
    b.build_filter(np.zeros((0, 3), np.int32))
    assert [len(x) for x in b.download_filter()] == [0, 1, 0]
    a.close(); b.close()


@pytest.mark.parametrize("D,batch", [(128, 2048), (256, 2048), (128, 48), (512, 2048), (200, 2048), (36, 7)])
def test_transe_compact_tables_many_rows(D, batch):
    """Candidates with up to 150 rows per step (beyond the rows whose indices are prefetched), several row widths
    (one and more float4 per lane, one and two rows in flight per warp): compact == full tables bit for bit, and both
    match the oracle on a small table."""
    from kelpie_b200 import plans, runtime
    rng = np.random.default_rng(D + batch)
    N, R = 700, 9
    ent = (rng.standard_normal((N, D)) * 0.2).astype(np.float32)
    rel = (rng.standard_normal((2 * R, D)) * 0.2).astype(np.float32)
    ctx = runtime.Context("TransE", ent, rel, norm=2)
    hp = dict(batch_size=batch, epochs=5, lr=0.01, margin=2.0, negative_triples_ratio=5, regularizer_weight=1.0)
    jobs = _jobs(rng, N, R, D, [75, 0, 1, 40, 13, 64, 3], 0.3)
    w = ko.Weights("TransE", torch.from_numpy(ent), torch.from_numpy(rel), norm=2, init_scale=1e-3)
    kg = ko.KG(np.zeros((0, 3), np.int64), np.zeros((0, 3), np.int64), np.zeros((0, 3), np.int64), N, R)
    seed_all(5)
    want = np.stack([ko.post_train(w, kg, torch.from_numpy(init).view(1, -1), facts, hp)[-1].numpy() for facts, init in jobs])
    seed_all(5)
    b = plans.Batch("TransE", N, R, hp)
    for facts, init in jobs:
        b.add(facts, init)
    full = ctx.post_train(runtime.make_hp("TransE", hp), **b.arrays()).cpu().numpy()
    compact = ctx.post_train(runtime.make_hp("TransE", hp), **b.arrays(compact=True)).cpu().numpy()
    _assert_rows(full, want)
    assert np.array_equal(full, compact)
    assert np.isfinite(full).all() and np.abs(full - np.stack(b.init_rows)).max() > 0
    ctx.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
@pytest.mark.parametrize("kind", ["TransE", "ComplEx"])
def test_two_devices_in_one_process(kind):
    """One context per GPU inside ONE process (a thread per context instead of a process per GPU): the kernels'
    shared-memory opt-in is a per-device attribute and must be set on every device, not once per process."""
    from kelpie_b200 import plans, runtime
    rng = np.random.default_rng(5)
    N, R, D = 3001, 6, 512 if kind == "ComplEx" else 256
    ent = (rng.standard_normal((N, D)) * 0.2).astype(np.float32)
    rel = (rng.standard_normal((2 * R, D)) * 0.2).astype(np.float32)
    hp = (dict(optimizer_name="Adagrad", batch_size=512, epochs=3, lr=0.043, decay1=0.9, decay2=0.999, regularizer_name="N3",
               regularizer_weight=0) if kind == "ComplEx" else
          dict(batch_size=2048, epochs=3, lr=0.01, margin=2.0, negative_triples_ratio=5, regularizer_weight=1.0))
    jobs = _jobs(rng, N, R, D, [5] * 300, 1e-3 if kind == "ComplEx" else 0.3)
    triples = np.tile(np.array([[N, 1, 7]], dtype=np.int32), (len(jobs), 1))
    off, ids = np.zeros(len(jobs) + 1, np.int64), np.zeros(1, np.int32)
    out = []
    for dev in (1, 0):  # the second device first: a per-process flag would leave it unconfigured
        seed_all(9)
        b = plans.Batch(kind, N, R, hp)
        for facts, init in jobs:
            b.add(facts, init)
        with torch.cuda.device(dev):
            ctx = runtime.Context(kind, ent, rel, norm=2, device=dev)
            rows = ctx.post_train(runtime.make_hp(kind, hp), **b.arrays(compact=True))
            mode = runtime.RANK_ENGINE_MIN if kind == "TransE" else runtime.RANK_ENGINE_MAX
            ts, bs, rk = ctx.filtered_rank(triples, mode, mimic_rows=rows, flt_off=off, flt_ids=ids)
            out.append((rows.cpu().numpy(), rk.cpu().numpy()))
            ctx.close()
    assert np.array_equal(out[0][1], out[1][1])
    np.testing.assert_allclose(out[0][0], out[1][0], rtol=1e-5, atol=1e-7)
