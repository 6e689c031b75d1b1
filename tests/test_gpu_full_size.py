"""BASELINE.json configs[4] size (1M entities x ComplEx dim 256 = 512 floats per row): parity through
size-independent properties (the oracle would need minutes per candidate here)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N, DIM, R = 1_000_000, 256, 512


@pytest.fixture(scope="module")
def ctx():
    from kelpie_b200 import runtime
    g = torch.Generator(device="cuda").manual_seed(42)
    ent = torch.randn(N, 2 * DIM, generator=g, device="cuda") * 0.1
    rel = torch.randn(2 * R, 2 * DIM, generator=g, device="cuda") * 0.1
    c = runtime.Context("ComplEx", ent, rel)
    yield c
    c.close()


@pytest.mark.parametrize("Q", [4, 70])  # streaming kernel / 64-query tile kernel
def test_fused_rank_equals_count_over_materialised_scores(ctx, Q):
    from kelpie_b200 import runtime
    rng = np.random.default_rng(Q)
    triples = np.stack([rng.integers(0, N, Q), rng.integers(0, 2 * R, Q), rng.integers(0, N, Q)], 1)
    sc = ctx.all_scores(triples)
    t = sc.gather(1, torch.as_tensor(triples[:, 2], device=sc.device).view(-1, 1))
    lens = rng.integers(0, 300, Q)
    off = np.zeros(Q + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    ids = np.concatenate([np.sort(rng.choice(N, n, replace=False)) for n in lens]).astype(np.int32)
    masked = sc.clone()
    for q in range(Q):  # model.py:50-54
        masked[q, torch.as_tensor(ids[off[q]:off[q + 1]], device=sc.device, dtype=torch.long)] = -1e6
        masked[q, triples[q, 2]] = t[q, 0]
    want = (masked >= t).sum(1)
    ts, bs, rk = ctx.filtered_rank(triples, runtime.RANK_MODEL, flt_off=off, flt_ids=ids)
    assert torch.equal(rk, want)  # bit-exact integer ranks
    torch.testing.assert_close(ts, t.view(-1), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(bs, masked.max(1).values, rtol=1e-5, atol=1e-7)


def test_tensor_core_pass_equals_cuda_core_pass_at_1m(ctx):
    """bf16x3 tcgen05 post-training vs the fp32 CUDA-core pass: mimic rows within 1e-4 (row max norm)."""
    from kelpie_b200 import plans, runtime
    hp = dict(optimizer_name="Adagrad", batch_size=512, epochs=2, lr=0.043, decay1=0.9, decay2=0.999,
              regularizer_name="N3", regularizer_weight=0)
    rng = np.random.default_rng(3)
    torch.manual_seed(0)
    b = plans.Batch("ComplEx", N, R, hp)
    for _ in range(12):
        T = int(rng.integers(4, 9))
        facts = [((N, int(rng.integers(0, R)), int(rng.integers(0, N))) if rng.random() < 0.5
                  else (int(rng.integers(0, N)), int(rng.integers(0, R)), N)) for _ in range(T)]
        b.add(facts, (rng.random(2 * DIM) * 1e-3).astype(np.float32))
    arrs = b.arrays()
    rows = {}
    for simt in (1, 0):
        ctx.set_option("force_simt", simt)
        rows[simt] = ctx.post_train(runtime.make_hp("ComplEx", hp), **arrs).cpu().numpy()
    ctx.set_option("force_simt", 0)
    scale = np.abs(rows[1]).max(axis=1, keepdims=True)
    assert (np.abs(rows[0] - rows[1]) / scale).max() < 1e-4
    assert np.abs(rows[0] - arrs["init_rows"]).max() > 1e-3  # the rows did move


def test_fused_pass_matches_fp64_at_1m(ctx):
    """The kernel that produces the bench number (S/V-specialised cluster pass, 180 rows = two query tiles) against an
    fp64 restatement over ALL 1 000 000 x 512 entities: l = sum_j exp(z_j - m), O = sum_j p_j E_j.  Same stated
    tolerance as tests/test_gpu_contract.py: 2e-5 + 4 * 2^-17 * max_j |q o E_j|_2, relative to sum p and sum p |E|."""
    rng = np.random.default_rng(5)
    q = torch.from_numpy((rng.standard_normal((180, 2 * DIM)) * 0.25).astype(np.float32)).cuda()
    m, l, O = ctx.contract(q, 0)
    torch.cuda.synchronize()
    q64 = q.double()
    zmax = torch.full((180,), -float("inf"), dtype=torch.float64, device="cuda")
    for j0 in range(0, N, 100_000):  # pass 1: the exact row maxima
        zmax = torch.maximum(zmax, (q64 @ ctx.ent[j0:j0 + 100_000].double().T).max(1).values)
    rl = torch.zeros(180, dtype=torch.float64, device="cuda")
    rO = torch.zeros(180, 2 * DIM, dtype=torch.float64, device="cuda")
    cond = torch.zeros_like(rO)
    zerr = torch.zeros(180, dtype=torch.float64, device="cuda")
    for j0 in range(0, N, 100_000):
        e = ctx.ent[j0:j0 + 100_000].double()
        p = torch.exp(q64 @ e.T - zmax[:, None])
        rl += p.sum(1)
        rO += p @ e
        cond += p @ e.abs()
        zerr = torch.maximum(zerr, torch.sqrt((q64 ** 2) @ (e ** 2).T).max(1).values)
    tol = 2e-5 + 4.0 * zerr * 2.0 ** -17
    scale = torch.exp(zmax - m.double())  # the kernel's lazy reference max may trail the true max by up to 8
    assert (m.double() <= zmax + 1e-4).all() and (m.double() >= zmax - 8.0 - 1e-3).all()
    assert ((l.double() / scale - rl).abs() <= tol * rl).all()
    assert (((O.double() / scale[:, None] - rO).abs()).max(1).values <= tol * cond.max(1).values).all()


def test_headline_config_matches_the_oracle(ctx):
    """BASELINE configs[4] against the ORACLE (not against another kernel of this repo): 3 candidates of 60 facts
    (180 mimic-lhs rows: two query tiles, i.e. the S/V-specialised cluster kernel that produces the bench number),
    3 epochs of Adagrad over all 1 000 000 x 512 entities, then the filtered rank.  Reference path:
    multiclass_nll_optimizer.py:123-164, complex.py:59-86, post_training_engine.py:101-125.
    Tolerances: post-trained mimic rows 1e-4 of the row's max |.| -- but not below the reference arithmetic's own
    reproducibility at this size: Adagrad's update lr * g / sqrt(sum g^2) is scale-invariant, so a component whose
    gradient is ~1e-3 of the typical one turns an fp32 summation difference over 1 000 000 entities straight into a
    row difference (measured on B200: after ONE epoch every path agrees with the oracle to 2e-7; after two, the
    fp32 oracle, its fp64 restatement, this repo's fp32 CUDA-core pass and the tcgen05 pass differ PAIRWISE by
    0.8 - 1.3e-4).  So the rows are held to max(1e-4, 3 x |fp32 oracle - fp64 oracle|) against the fp64 oracle, the
    noise floor is printed, and the one-epoch rows are held to 1e-5.  Target score: 1e-4 of the score range; integer
    rank exact away from ties (see the comments at the assertion)."""
    from oracle import kelpie_oracle as ko
    from kelpie_b200 import plans, runtime
    hp = dict(optimizer_name="Adagrad", batch_size=512, epochs=3, lr=0.043, decay1=0.9, decay2=0.999,
              regularizer_name="N3", regularizer_weight=0)
    rng = np.random.default_rng(11)
    torch.manual_seed(0)
    ent_h, rel_h = ctx.ent.cpu(), ctx.rel.cpu()
    w = ko.Weights("ComplEx", ent_h, rel_h, init_scale=1e-3)
    kg = ko.KG(np.zeros((0, 3), np.int64), np.zeros((0, 3), np.int64), np.zeros((0, 3), np.int64), N, R)
    p, o = int(rng.integers(0, R)), int(rng.integers(0, N))
    jobs, inits, filters = [], [], []
    b = plans.Batch("ComplEx", N, R, hp)
    for _ in range(3):
        x, r = rng.choice(N, 60, replace=False), rng.integers(0, R, 60)
        facts = np.stack([np.full(60, N), r, x], 1)  # every fact has the mimic as head: 60 rows take the full pass
        init = (rng.random(2 * DIM) * 1e-3).astype(np.float32)
        b.add(facts, init)
        jobs.append(facts)
        inits.append(init)
        filters.append(np.unique(np.concatenate([rng.integers(0, N, 5), [o]])).astype(np.int32))
    arrs = b.arrays()
    assert int((arrs["pos"][:, 0] == N).sum()) == 180
    ctx.set_option("timing", 1)
    ctx.stat("reset")
    rows = ctx.post_train(runtime.make_hp("ComplEx", hp), **arrs)
    assert ctx.stat("n_flash") >= 3  # the fused tcgen05 pass ran once per epoch
    ctx.set_option("timing", 0)
    off = np.zeros(4, dtype=np.int64)
    off[1:] = np.cumsum([len(f) for f in filters])
    ts, bs, rk = ctx.filtered_rank(np.array([[N, p, o]] * 3, np.int32), runtime.RANK_ENGINE_MAX, mimic_rows=rows,
                                   flt_off=off, flt_ids=np.concatenate(filters).astype(np.int32))
    rows, ts, rk = rows.cpu().numpy(), ts.cpu().numpy(), rk.cpu().numpy()
    hp1 = dict(hp, epochs=1)
    b1 = plans.Batch("ComplEx", N, R, hp1)
    for f, i in zip(jobs, inits):
        b1.add(f, i)
    rows1 = ctx.post_train(runtime.make_hp("ComplEx", hp1), **b1.arrays()).cpu().numpy()
    w64 = ko.Weights("ComplEx", ent_h.double(), rel_h.double(), init_scale=1e-3)
    for c in range(3):
        init = torch.from_numpy(inits[c]).view(1, -1)
        one = ko.post_train(w, kg, init, jobs[c], hp1)[-1].numpy()
        assert np.abs(rows1[c] - one).max() <= 1e-5 * np.abs(one).max()  # one epoch: insensitive to summation order
        table = ko.post_train(w, kg, init, jobs[c], hp)
        want_row = table[-1].numpy()
        row64 = ko.post_train(w64, kg, init.double(), jobs[c], hp)[-1].numpy()
        scale_r = np.abs(row64).max()
        floor = np.abs(want_row - row64).max() / scale_r          # the fp32 reference arithmetic against its fp64 restatement
        err64 = np.abs(rows[c] - row64).max() / scale_r
        print(f"candidate {c}: rows vs fp64 oracle {err64:.2e}, fp32 oracle vs fp64 oracle {floor:.2e}, vs fp32 oracle "
              f"{np.abs(rows[c] - want_row).max() / scale_r:.2e}")
        assert err64 <= max(1e-4, 3.0 * floor)
        assert np.abs(want_row - inits[c]).max() > 1e-3  # the row did move
        res = ko.triple_results(w, table, (N, p, o), filters[c])
        with torch.no_grad():
            sc = ko.all_scores(w, table, np.array([[N, p, o]]))[0].numpy()
        scale = float(np.abs(sc).max())  # logits here are ~1e-3: tolerances are relative to the score range, not to 1
        assert abs(float(ts[c]) - res["target_score"]) <= 1e-4 * scale
        # Rank: exact away from ties.  Two legitimate sources of a difference: (i) entities that change side of the target
        # between our row and the oracle's row, both scored by the same device kernel; (ii) entities whose oracle score is
        # within fp32 summation round-off (4e-6 of the score range) of the target -- the CPU matmul and the device's FMA
        # chain add 512 products in different orders.
        q = np.array([[N, p, o]] * 2, np.int32)
        dev = ctx.all_scores(q, mimic_rows=np.stack([rows[c], want_row]).astype(np.float32)).cpu().numpy()
        flips = int(((dev[0] >= dev[0, o]) != (dev[1] >= dev[1, o])).sum())
        ties = int((np.abs(sc - res["target_score"]) <= 4e-6 * scale).sum()) - 1
        assert abs(int(rk[c]) - res["target_rank"]) <= flips + ties, (int(rk[c]), res["target_rank"], flips, ties)
        print(f"candidate {c}: rank {int(rk[c])} vs oracle {res['target_rank']} (side changes {flips}, round-off ties {ties}), "
              f"row err {np.abs(rows[c] - want_row).max() / np.abs(want_row).max():.2e}")


def test_rows_do_not_depend_on_the_batch(ctx):
    """A candidate's post-trained row must not depend on how many other candidates travel with it.  It used to: the tensor
    core's fp32 accumulation truncates, so a running sum in TMEM loses ~2^-24 of itself per accumulation -- 1.9e-3 over the
    31 000 K-steps of a 500 000-entity strip -- and large batches (few, long strips) drifted from small ones (many short
    strips) by up to 1.3e-3 of the row after 43 Adagrad epochs.  Strips are now at most 256 entity tiles long
    (kp_set_option "umma_max_tps") and merged in fp32: the same candidate in a batch of 2 and in a batch of 600 (several
    waves of clusters) agrees to 5e-5; without the bound the test's own measurement shows the drift."""
    from kelpie_b200 import plans, runtime
    hp = dict(optimizer_name="Adagrad", batch_size=512, epochs=43, lr=0.043, decay1=0.9, decay2=0.999,
              regularizer_name="N3", regularizer_weight=0)
    rng = np.random.default_rng(21)
    torch.manual_seed(0)
    jobs = []
    for _ in range(600):
        T = int(rng.integers(8, 65))
        x, r, head = rng.choice(N, T, replace=False), rng.integers(0, R, T), rng.random(T) < 0.5
        facts = np.where(head[:, None], np.stack([np.full(T, N), r, x], 1), np.stack([x, r, np.full(T, N)], 1))
        jobs.append((facts, (rng.random(2 * DIM) * 1e-3).astype(np.float32)))

    def run(n, max_tps):
        ctx.set_option("umma_max_tps", max_tps)
        b = plans.Batch("ComplEx", N, R, hp)
        for f, i in jobs[:n]:
            b.add(f, i)
        rows = ctx.post_train(runtime.make_hp("ComplEx", hp), **b.arrays()).cpu().numpy()[:2].astype(np.float64)
        ctx.set_option("umma_max_tps", 256)
        return rows

    small, big, big_unbounded = run(2, 256), run(600, 256), run(600, 0)
    scale = np.abs(small).max()
    drift, drift_unbounded = np.abs(big - small).max() / scale, np.abs(big_unbounded - small).max() / scale
    print(f"batch of 600 vs batch of 2: {drift:.2e} (strips <= 256 tiles), {drift_unbounded:.2e} (unbounded strips)")
    assert drift <= 5e-5
    assert drift_unbounded > drift  # the effect the bound removes
