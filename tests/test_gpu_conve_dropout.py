"""ConvE post-training with dropout > 0: the kernels' counter-based masks (kp_dropout.cuh) are
re-generated here in numpy and fed to a torch-autograd restatement of the reference step
(conve.py:133-158, bce_optimizer.py:161-208); mimic rows must agree within 1e-4."""
from collections import defaultdict

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.golden_util import load

pytestmark = pytest.mark.gpu
M64 = (1 << 64) - 1


def drop_scale(seed, pair, step, elem, p):
    """numpy port of kp_drop_scale (vectorised over elem)."""
    elem = np.asarray(elem, dtype=np.uint64)
    x = np.uint64(seed) ^ ((np.uint64(pair) << np.uint64(32)) | np.uint64(step))
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15) * (elem + np.uint64(1))
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    x = x ^ (x >> np.uint64(31))
    h = (x >> np.uint64(32)).astype(np.uint64)
    thr = np.uint64(int(p * 4294967296.0))
    return np.where(h < thr, 0.0, 1.0 / (1.0 - p)).astype(np.float32)


def features(c, lhs, rel, pair_ids, step, seed, rates):
    D = lhs.shape[1]
    H = D // 20
    x = torch.cat([lhs.view(-1, 1, 20, H), rel.view(-1, 1, 20, H)], 2)
    x = F.batch_norm(x, c["bn1_mean"], c["bn1_var"], c["bn1_w"], c["bn1_b"], False, 0.1, 1e-5)
    if rates[0] > 0:
        m = np.stack([drop_scale(seed, p, step, np.arange(40 * H), rates[0]) for p in pair_ids])
        x = x * torch.from_numpy(m).view(-1, 1, 40, H)
    x = F.conv2d(x, c["conv_w"], c["conv_b"])
    x = torch.relu(F.batch_norm(x, c["bn2_mean"], c["bn2_var"], c["bn2_w"], c["bn2_b"], False, 0.1, 1e-5))
    if rates[1] > 0:
        m = np.stack([drop_scale(seed, p, step, (1 << 20) + np.arange(x.shape[1]), rates[1]) for p in pair_ids])
        x = x * torch.from_numpy(m).view(-1, x.shape[1], 1, 1)
    x = F.linear(x.view(x.shape[0], -1), c["fc_w"], c["fc_b"])
    if rates[2] > 0:
        m = np.stack([drop_scale(seed, p, step, (2 << 20) + np.arange(D), rates[2]) for p in pair_ids])
        x = x * torch.from_numpy(m)
    return torch.relu(F.batch_norm(x, c["bn3_mean"], c["bn3_var"], c["bn3_w"], c["bn3_b"], False, 0.1, 1e-5))


def torch_post_train(w, ent, rel_t, facts, init, hp, R, seed, rates, pair_base):
    N = ent.shape[0]
    param = torch.nn.Parameter(torch.from_numpy(init).view(1, -1).clone())
    opt = torch.optim.Adam([param])
    f = np.asarray(facts, dtype=np.int64)
    inv = f.copy(); inv[:, 0], inv[:, 2] = f[:, 2], f[:, 0]; inv[:, 1] += R
    vocab = defaultdict(list)
    for s, p, o in np.vstack((f, inv)):
        vocab[(int(s), int(p))].append(int(o))
    pairs = list(vocab)
    bs, ls, step = hp["batch_size"], hp["label_smoothing"], 0
    for _ in range(hp["epochs"]):
        for b0 in range(0, len(pairs), bs):
            batch = pairs[b0:b0 + bs]
            table = torch.cat([ent, param], 0)
            targets = torch.zeros(len(batch), N + 1)
            for i, pr in enumerate(batch):
                targets[i, vocab[pr]] = 1.0
            targets = (1.0 - ls) * targets + 1.0 / (N + 1)
            lhs = table[[s for s, _ in batch]]
            rl = rel_t[[p for _, p in batch]]
            ids = [pair_base + b0 + i for i in range(len(batch))]
            x = features(w.conve, lhs, rl, ids, step, seed, rates)
            pred = torch.sigmoid(x @ table.t())
            opt.zero_grad()
            F.binary_cross_entropy(pred, targets).backward()
            opt.step()
            step += 1
    return param.detach().numpy().reshape(-1)


@pytest.mark.parametrize("rates", [(0.0, 0.0, 0.2), (0.1, 0.25, 0.3)])
def test_conve_dropout_matches_torch_with_same_masks(rates):
    from kelpie_b200 import plans, runtime
    z, meta, kg, w, order = load("ConvE")
    N, R, D = kg.num_entities, kg.num_relations, w.dim
    conve = dict(w.conve)
    conve["dropout"] = rates
    ctx = runtime.Context("ConvE", z["w_ent"], z["w_rel"], conve=conve)
    hp = dict(meta["hp"], batch_size=4, epochs=3)
    rng = np.random.default_rng(21)
    jobs = []
    for T in (3, 5):
        facts = [((N, int(rng.integers(0, R)), int(rng.integers(0, N))) if rng.random() < 0.5
                  else (int(rng.integers(0, N)), int(rng.integers(0, R)), N)) for _ in range(T)]
        jobs.append((facts, rng.random(D).astype(np.float32)))
    b = plans.Batch("ConvE", N, R, hp)
    for f, i in jobs:
        b.add(f, i)
    arrs = b.arrays()
    seed = 0x1234ABCD5678
    got = ctx.post_train(runtime.make_hp("ConvE", hp), dropout_seed=seed, **arrs).cpu().numpy()
    ent, rel_t = torch.from_numpy(z["w_ent"].copy()), torch.from_numpy(z["w_rel"].copy())
    for j, (f, i) in enumerate(jobs):
        want = torch_post_train(w, ent, rel_t, f, i, hp, R, seed, rates, int(arrs["row_off"][j]))
        assert np.abs(got[j] - want).max() <= 1e-4 * np.abs(want).max()
    # and the masks matter: a different seed gives different rows
    other = ctx.post_train(runtime.make_hp("ConvE", hp), dropout_seed=seed + 1, **arrs).cpu().numpy()
    assert np.abs(other - got).max() > 1e-6


@pytest.mark.parametrize("rates", [(0.0, 0.0, 0.0), (0.1, 0.25, 0.3)])
def test_conv_kernel_split_output_is_the_split_pass(rates):
    """>= 128 pairs per step: the Linear layer runs on tcgen05 and the conv kernel writes its bf16 hi / lo operand itself
    (kp_conve.cu); the rows must be those of the fp32 feature maps followed by the separate split pass, bit for bit,
    and stay within 1e-4 of the CUDA-core Linear layer."""
    from kelpie_b200 import plans, runtime
    z, meta, kg, w, order = load("ConvE")
    N, R, D = kg.num_entities, kg.num_relations, w.dim
    conve = dict(w.conve)
    conve["dropout"] = rates
    hp = dict(meta["hp"], batch_size=8, epochs=3)
    rng = np.random.default_rng(5)
    b = plans.Batch("ConvE", N, R, hp)
    for _ in range(70):  # ~ 70 * 6 pairs per step, both directions
        T = int(rng.integers(4, 9))
        facts = [((N, int(rng.integers(0, R)), int(rng.integers(0, N))) if rng.random() < 0.5
                  else (int(rng.integers(0, N)), int(rng.integers(0, R)), N)) for _ in range(T)]
        b.add(facts, rng.random(D).astype(np.float32))
    arrs = b.arrays()
    rows = {}
    # gemm_ksplit (default on) cuts the forward Linear layer's K over several CTA pairs when there are few rows and adds the partial
    # results in a fixed order: other roundings than one accumulation chain, but the same bits every run
    for name, opts in (("split", {"gemm_ksplit": 0}), ("fp32", {"conv_split": 0, "gemm_ksplit": 0}), ("simt", {"umma_fc": 0}),
                       ("ksplit", {}), ("ksplit_again", {})):
        ctx = runtime.Context("ConvE", z["w_ent"], z["w_rel"], conve=conve)
        for k, v in opts.items():
            ctx.set_option(k, v)
        rows[name] = ctx.post_train(runtime.make_hp("ConvE", hp), dropout_seed=99, **arrs).cpu().numpy()
    assert np.isfinite(rows["split"]).all()
    assert np.array_equal(rows["split"], rows["fp32"])
    assert np.abs(rows["split"] - rows["simt"]).max() <= 1e-4 * np.abs(rows["simt"]).max()
    assert np.array_equal(rows["ksplit"], rows["ksplit_again"])
    assert np.abs(rows["ksplit"] - rows["simt"]).max() <= 1e-4 * np.abs(rows["simt"]).max()
