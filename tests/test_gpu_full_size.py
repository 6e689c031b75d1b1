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
