"""Filtered rank on the tensor cores (kp_rank_umma.cu) against the exact fp32 pass (kp_pass.cu): the
counters (strictly better / tied / tied with a smaller id) and the ranks must be IDENTICAL -- the
tcgen05 pass only decides pairs whose score is further from the target than its error margin and
hands the rest to an exact re-check.  Covers ComplEx (no activation) and ConvE (sigmoid, incl. a
saturated regime where most pairs tie), duplicate entity rows (exact ties), targets inside their own
filter list, mimic rows, a table that does not fill its last tile, and 1M x 512 at full size."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _queries(rng, Q, N, R2, mimic_every=0):
    t = np.stack([rng.integers(0, N, Q), rng.integers(0, R2, Q), rng.integers(0, N, Q)], axis=1).astype(np.int32)
    if mimic_every:
        t[::mimic_every, 0] = N
    return t


def _filters(rng, Q, N, tgt, mean=6):
    off, ids = [0], []
    for q in range(Q):
        n = int(rng.poisson(mean))
        f = set(int(x) for x in rng.integers(0, N, n))
        if q % 3 == 0:
            f.add(int(tgt[q]))  # the target itself is a known fact (necessary mode)
        ids.extend(sorted(f))
        off.append(len(ids))
    return np.asarray(off, np.int64), np.asarray(ids if ids else [0], np.int32)


def _both(ctx, triples, mode, mimic, off, ids):
    out = {}
    for opt in (1, 0):
        ctx.set_option("umma_rank", opt)
        before = ctx.stat("rank_rechecks")
        ts, bs, rk, cn = ctx.filtered_rank(triples, mode, mimic_rows=mimic, flt_off=off, flt_ids=ids, counters=True)
        torch.cuda.synchronize()
        out[opt] = (ts.cpu().numpy(), bs.cpu().numpy(), rk.cpu().numpy(), cn.cpu().numpy(), ctx.stat("rank_rechecks") - before)
    ctx.set_option("umma_rank", 1)
    return out


def _assert_same(out, Q, N, max_recheck_frac, best_tol=None):
    ts1, bs1, rk1, cn1, n1 = out[1]
    ts0, bs0, rk0, cn0, n0 = out[0]
    assert n0 == 0
    assert np.array_equal(cn1, cn0)
    assert np.array_equal(rk1, rk0)
    assert np.array_equal(ts1, ts0)
    fin = np.isfinite(bs0)
    assert np.array_equal(np.isfinite(bs1), fin)
    if best_tol is None:
        assert np.abs(bs1[fin] - bs0[fin]).max() <= 1e-4 * max(1.0, np.abs(bs0[fin]).max())
    else:  # L2: the tensor-core estimate of the best distance is accurate in the SQUARED domain
        assert np.abs(bs1[fin].astype(np.float64) ** 2 - bs0[fin].astype(np.float64) ** 2).max() <= best_tol
    assert n1 <= max_recheck_frac * Q * N, f"{n1} re-checks for {Q} x {N} pairs"


@pytest.mark.parametrize("D,N,Q", [(400, 5003, 700), (512, 9001, 300), (128, 3000, 129)])
@pytest.mark.parametrize("mode", [1, 2])
def test_complex_rank_identical(D, N, Q, mode):
    from kelpie_b200 import runtime
    rng = np.random.default_rng(D + N + Q + mode)
    ent = (rng.standard_normal((N, D)) * 0.3).astype(np.float32)
    ent[N // 2: N // 2 + 40] = ent[7:47]  # duplicate rows: exact ties in fp32 and on the tensor cores
    rel = (rng.standard_normal((14, D)) * 0.3).astype(np.float32)
    ctx = runtime.Context("ComplEx", ent, rel)
    triples = _queries(rng, Q, N, 14, mimic_every=5)
    triples[1::7, 2] = 10  # targets among the duplicated rows
    mimic = (rng.standard_normal((Q, D)) * 0.3).astype(np.float32)
    off, ids = _filters(rng, Q, N, triples[:, 2])
    out = _both(ctx, triples, mode, mimic, off, ids)
    _assert_same(out, Q, N, 0.05)
    assert out[1][3][:, 1].sum() > 0  # ties exist
    ctx.close()


@pytest.mark.parametrize("scale", [1.0, 40.0])
def test_conve_rank_identical(scale):
    """scale = 40: logits far into the sigmoid's saturation, most entities tie with the target."""
    from kelpie_b200 import runtime
    from tests.golden_util import load
    z, meta, kg, w, order = load("ConvE")
    rng = np.random.default_rng(3)
    N, R2, D = int(z["n_ent"]), 2 * int(z["n_rel"]), int(z["w_ent"].shape[1])
    ent = (z["w_ent"] * scale).astype(np.float32)
    ctx = runtime.Context("ConvE", ent, z["w_rel"], conve=dict(w.conve))
    Q = 260
    triples = _queries(rng, Q, N, R2, mimic_every=4)
    mimic = rng.random((Q, D)).astype(np.float32)
    off, ids = _filters(rng, Q, N, triples[:, 2], mean=3)
    for mode in (1, 3):
        out = _both(ctx, triples, mode, mimic, off, ids)
        _assert_same(out, Q, N, 1.0)
    ctx.close()


def test_rank_full_size_matches_exact_pass():
    """Config 5 shape (1M x 512): 384 queries; also the size-independent property rank(target row = best) = 1."""
    from kelpie_b200 import runtime
    torch.manual_seed(0)
    N, D, Q = 1_000_000, 512, 384
    ent = torch.randn(N, D, device="cuda") * 0.1
    rel = torch.randn(8, D, device="cuda") * 0.1
    ctx = runtime.Context("ComplEx", ent, rel)
    rng = np.random.default_rng(1)
    triples = _queries(rng, Q, N, 8)
    off, ids = _filters(rng, Q, N, triples[:, 2])
    out = _both(ctx, triples, 2, None, off, ids)
    _assert_same(out, Q, N, 0.01)
    ctx.close()


@pytest.mark.parametrize("D,N,Q", [(256, 5003, 700), (128, 9001, 300), (200, 3000, 129)])
@pytest.mark.parametrize("mode", [0, 2])
def test_transe_l2_rank_identical(D, N, Q, mode):
    """TransE with the L2 norm (a minimiser: transe.py:48-65) through the tensor-core pass: squared-distance
    margins, exact re-check with the exact pass's chain.  Duplicate rows (exact ties), near-duplicates (distances
    that differ in the last bits), targets at distance ~0 and mimic heads."""
    from kelpie_b200 import runtime
    rng = np.random.default_rng(D + N + Q + mode)
    ent = (rng.standard_normal((N, D)) * 0.3).astype(np.float32)
    ent[N // 2: N // 2 + 40] = ent[7:47]                                             # exact ties
    ent[N // 3: N // 3 + 40] = ent[7:47] * np.float32(1 + 2e-7)                      # near ties
    rel = (rng.standard_normal((14, D)) * 0.3).astype(np.float32)
    rel[3] = 0.0                                                                     # (h, 3, h): distance exactly 0
    ctx = runtime.Context("TransE", ent, rel, norm=2)
    triples = _queries(rng, Q, N, 14, mimic_every=5)
    triples[1::7, 2] = 10  # targets among the duplicated rows
    triples[2::11, 1] = 3
    triples[2::11, 2] = np.where(triples[2::11, 0] == N, 5, triples[2::11, 0])
    mimic = (rng.standard_normal((Q, D)) * 0.3).astype(np.float32)
    off, ids = _filters(rng, Q, N, triples[:, 2])
    out = _both(ctx, triples, mode, mimic, off, ids)
    _assert_same(out, Q, N, 0.05, best_tol=2e-3 * 2 * D * 0.09)
    assert out[1][4] > 0          # the tensor-core pass ran (it reports its re-checks)
    assert out[1][3][:, 1].sum() > 0  # ties exist
    ctx.close()


def test_transe_l2_rank_dbpedia50_shape():
    """BASELINE configs[0] shape (24 620 x 256, 4097 queries with their own mimic head): ranks identical, few re-checks."""
    from kelpie_b200 import runtime
    rng = np.random.default_rng(11)
    N, D, Q = 24620, 256, 4097
    ent = (rng.standard_normal((N, D)) * (2.0 / (N + D)) ** 0.5 * 30).astype(np.float32)
    rel = (rng.standard_normal((702, D)) * 0.05).astype(np.float32)
    ctx = runtime.Context("TransE", ent, rel, norm=2)
    triples = _queries(rng, Q, N, 702, mimic_every=1)
    mimic = (rng.standard_normal((Q, D)) * 0.1).astype(np.float32)
    off, ids = _filters(rng, Q, N, triples[:, 2])
    out = _both(ctx, triples, 0, mimic, off, ids)
    _assert_same(out, Q, N, 0.02, best_tol=2e-2)
    ctx.close()
