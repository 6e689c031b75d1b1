"""The tcgen05 fused pass alone (kp_debug_contract) against an fp64 restatement of SURVEY 9.3/9.4's
inner contraction: l = sum_j exp(z_gj - m), O = sum_j p_gj E[j] (softmax) or p = sigmoid(z).
Stated tolerance per row: (2e-5 + 4 * 2^-17 * max_j |q o E_j|_2) of sum_j p_gj (l) and of
sum_j p_gj |E_jk| (O) -- the bf16x3 split keeps ~16 mantissa bits per product (fp32 accumulation), and
an absolute perturbation of a logit is a relative perturbation of its probability.

Covers the cluster-of-4 kernels (rows wider than 256 floats: the default one, where one SM pair
scores and the other contracts, and round 1's, where both alternate; P crosses distributed shared
memory in both), the cta_group::2 pair kernel and the single-CTA kernel, with
strips of many entity tiles, a ragged last tile, and logits that grow along the table so that the
lazy rescale of the TMEM accumulators fires (reference max moves by more than 8)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(q, ent, mode):
    z = q.astype(np.float64) @ ent.astype(np.float64).T
    if mode == 0:
        m = z.max(axis=1)
        p = np.exp(z - m[:, None])
    else:
        m = np.zeros(len(q))
        p = 1.0 / (1.0 + np.exp(-z))
    e = ent.astype(np.float64)
    # rounding of the split products perturbs z by ~2^-17 |q o E_j|_2, i.e. p by that relative amount
    zerr = np.sqrt((q.astype(np.float64) ** 2) @ (e ** 2).T).max(axis=1) * 2.0 ** -17
    return m, p.sum(axis=1), p @ e, p @ np.abs(e), zerr


def _check(ctx, q, ent, mode, tol=2e-5):
    m, l, O = ctx.contract(q, mode)
    torch.cuda.synchronize()
    m, l, O = m.cpu().numpy().astype(np.float64), l.cpu().numpy().astype(np.float64), O.cpu().numpy().astype(np.float64)
    rm, rl, rO, cond, zerr = _reference(q, ent, mode)
    tol = tol + 4.0 * zerr
    if mode == 0:
        assert np.all(m <= rm + 1e-4 * np.abs(rm).max() + 1e-6) and np.all(m >= rm - 8.0 - 1e-3)
        scale = np.exp(rm - m)  # the kernel's reference max may trail the true max by up to 8
        l, O = l / scale, O / scale[:, None]
    assert (np.abs(l - rl) <= tol * np.abs(rl)).all()
    # forward-error bound of the contraction: relative to sum_j p_j |E_jk| (O itself cancels when p is flat)
    assert (np.abs(O - rO).max(axis=1) <= tol * cond.max(axis=1)).all()


@pytest.mark.parametrize("D,N,G", [(512, 20011, 300), (400, 9001, 1400), (512, 3001, 260), (256, 9001, 300), (128, 5000, 100), (272, 5003, 300), (304, 5003, 520)])
@pytest.mark.parametrize("mode", [0, 1])
def test_contract_matches_fp64(D, N, G, mode):
    from kelpie_b200 import runtime
    rng = np.random.default_rng(D + N + G + mode)
    ent = (rng.standard_normal((N, D)) * 0.3).astype(np.float32)
    rel = np.zeros((2, D), np.float32)
    q = (rng.standard_normal((G, D)) * (0.25 if mode == 0 else 0.05)).astype(np.float32)
    ctx = runtime.Context("ComplEx", ent, rel)
    _check(ctx, q, ent, mode)
    ctx.close()


@pytest.mark.parametrize("x4", [2, 1, 0])
def test_contract_rescale_path(x4):
    """Entity norms grow along the table -> the row max keeps moving up -> O is rescaled in TMEM, by
    the tile's owner and (cluster-of-4 kernel) by the pair that received the tile."""
    from kelpie_b200 import runtime
    rng = np.random.default_rng(7)
    N, D, G = 12001, 512, 256
    ent = rng.standard_normal((N, D)).astype(np.float32) * 0.2
    ent *= np.linspace(0.2, 3.0, N, dtype=np.float32)[:, None]
    q = (rng.standard_normal((G, D)) * 0.6).astype(np.float32)
    z = q.astype(np.float64) @ ent.astype(np.float64).T
    assert (z.max(axis=1) - z[:, :128].max(axis=1)).min() > 16.0  # several rescales per row
    ctx = runtime.Context("ComplEx", ent, np.zeros((2, D), np.float32))
    ctx.set_option("umma_x4", x4)
    _check(ctx, q, ent, 0)
    ctx.close()


@pytest.mark.parametrize("quad", [2, 1])
def test_quad_matches_pair_kernel(quad):
    from kelpie_b200 import runtime
    rng = np.random.default_rng(11)
    N, D, G = 30011, 512, 512
    ent = (rng.standard_normal((N, D)) * 0.3).astype(np.float32)
    q = (rng.standard_normal((G, D)) * 0.3).astype(np.float32)
    ctx = runtime.Context("ComplEx", ent, np.zeros((2, D), np.float32))
    out = {}
    for x4 in (quad, 0):
        ctx.set_option("umma_x4", x4)
        m, l, O = ctx.contract(q, 0)
        torch.cuda.synchronize()
        out[x4] = (m.cpu().numpy(), l.cpu().numpy(), O.cpu().numpy())
    s = np.exp(out[quad][0].astype(np.float64) - out[0][0])
    assert np.abs(out[quad][1] * s - out[0][1]).max() <= 1e-5 * np.abs(out[0][1]).max()
    assert np.abs(out[quad][2] * s[:, None] - out[0][2]).max() <= 1e-5 * np.abs(out[0][2]).max()
    ctx.close()


@pytest.mark.parametrize("N,G", [(100, 300), (130, 257), (255, 512), (257, 300), (1000, 1), (77, 130)])
def test_contract_tiny_tables_and_ragged_rows(N, G):
    """S/V kernel edge cases: fewer entities than one tile pair (the second tile of the pair is all padding), a table that
    ends just past a tile boundary, row counts that are not a multiple of the 256-row cluster tile, a single row."""
    from kelpie_b200 import runtime
    rng = np.random.default_rng(N * 1000 + G)
    D = 512
    ent = (rng.standard_normal((N, D)) * 0.3).astype(np.float32)
    q = (rng.standard_normal((G, D)) * 0.25).astype(np.float32)
    ctx = runtime.Context("ComplEx", ent, np.zeros((2, D), np.float32))
    for mode in (0, 1):
        _check(ctx, q if mode == 0 else q * 0.2, ent, mode)
    ctx.close()
