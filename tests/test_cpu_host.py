"""CPU-only checks: the C-ABI library loads and exports every declared symbol, the host-side
mirrors (dataset overlay, batch plans, filter CSR) agree with the oracle, and the candidate
sharding works across two gloo ranks.  No kernel is launched here."""
import ctypes
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import kelpie_oracle as ko
from tests.golden_util import load, seed_all

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from kelpie_b200 import runtime
    lib = runtime.load_library()
    header = open(os.path.join(ROOT, "include", "kelpie_b200.h")).read()
    declared = set(re.findall(r"\b(kp_[a-z_0-9]+)\s*\(", header))
    assert declared == set(runtime.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.kp_abi_version() == 3


def test_no_cpu_fallback():
    """Without a CUDA device context creation fails loudly with KP_ECUDA (no CPU path)."""
    from kelpie_b200 import runtime
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    lib = runtime.load_library()
    ent = np.zeros((8, 8), np.float32)
    h = ctypes.c_void_p()
    rc = lib.kp_ctx_create(0, 0, 8, 4, 8, 2, ent.ctypes.data, ent.ctypes.data, None, ctypes.byref(h))
    assert rc == -2 and b"no CUDA device" in lib.kp_last_error(None)
    with pytest.raises(RuntimeError):
        runtime.Context("TransE", ent, ent)


def test_struct_layouts_match_header():
    from kelpie_b200 import runtime
    assert ctypes.sizeof(runtime.HP) == 44  # ABI 3: + regularizer
    assert ctypes.sizeof(runtime.PTBatch) == 96 + 4 * 8  # ABI 2: the compact TransE tables
    assert ctypes.sizeof(runtime.ConvEWeights) == 7 * 8 + 5 * 4 + 4


@pytest.mark.parametrize("kind", ["TransE", "ComplEx", "ConvE"])
def test_batch_plans_match_oracle_draws(kind):
    """plans.Batch consumes the generators exactly like the oracle (hence the reference)."""
    from kelpie_b200 import plans
    z, meta, kg, w, order = load(kind)
    hp = meta["hp"]
    N, R = kg.num_entities, kg.num_relations
    e = next(iter(order))
    facts = [ko._swap(t, e, N) for t in order[e]]
    seed_all(7)
    log = []
    ko.post_train(w, kg, torch.rand(1, w.dim), facts, hp, log)
    seed_all(7)
    torch.rand(1, w.dim)
    b = plans.Batch(kind, N, R, hp)
    b.add(facts, np.zeros(w.dim, np.float32))
    a = b.arrays()
    if kind == "TransE":
        pos = np.concatenate([s["pos"] for s in log])
        neg = np.concatenate([s["neg"] for s in log])
        np.testing.assert_array_equal(a["pos"], pos)
        np.testing.assert_array_equal(a["neg"], neg)
        c = b.arrays(compact=True)  # the 6-byte index rows describe the same tables
        assert c["pos"] is None and c["pos_idx"].dtype == np.uint16 and c["neg_code"].dtype == np.int32
        cp, cn = plans.expand_transe(c["facts"][c["fact_off"][0]:c["fact_off"][1]], c["pos_idx"], c["neg_code"])
        np.testing.assert_array_equal(cp, pos)
        np.testing.assert_array_equal(cn, neg)
    elif kind == "ComplEx":
        assert a["static_epochs"]
        got = sorted(map(tuple, a["pos"]))
        for s in log:  # every step covers the same multiset of rows
            assert sorted(map(tuple, s["rows"])) == got
        assert torch.rand(1).item() == pytest.approx(_after(kind, w, kg, facts, hp))
    else:
        np.testing.assert_array_equal(a["pos"][:, :2], log[0]["pairs"])


def _after(kind, w, kg, facts, hp):
    seed_all(7)
    ko.post_train(w, kg, torch.rand(1, w.dim), facts, hp)
    return torch.rand(1).item()


def test_kelpie_dataset_overlay_matches_oracle_mimic():
    from kelpie_b200.data import Dataset, KelpieDataset
    z, meta, kg, w, order = load("TransE")
    ds = Dataset("g", z["train"], z["valid"], z["test"], kg.num_entities, kg.num_relations)
    e = next(iter(order))
    ds.entity_to_training_triples[e] = order[e]
    kd = KelpieDataset(ds, e)
    mim = ko.Mimic(kg, e, order[e])
    assert [tuple(t) for t in kd.kelpie_training_triples] == mim.base_facts
    rule = [order[e][0], order[e][2]]
    kd.remove_training_triples(rule)
    facts, flt = mim.without(rule)
    assert [tuple(t) for t in kd.kelpie_training_triples] == facts
    for key, v in flt.items():
        if key[0] == mim.M:
            assert sorted(kd.to_filter[key]) == sorted(v)
    kd.undo_removal()
    assert [tuple(t) for t in kd.kelpie_training_triples] == mim.base_facts
    for key, v in mim.filter.items():
        if key[0] == mim.M:
            assert sorted(kd.to_filter[key]) == sorted(v)
    kd.add_training_triples(rule[:1])
    facts, flt = mim.with_added(rule[:1])
    assert [tuple(t) for t in kd.kelpie_training_triples] == facts
    kd.undo_addition()
    # the dataset's own dicts were never touched
    assert all(kg.to_filter[k] == list(ds.to_filter[k]) for k in kg.to_filter)


def test_filter_csr_is_sorted_and_deduplicated():
    from kelpie_b200.runtime import filter_csr
    keys, off, ids = filter_csr({(3, 1): [5, 2, 5], (0, 2): [7], (3, 0): []}, 4)
    np.testing.assert_array_equal(keys, [2, 13])
    np.testing.assert_array_equal(off, [0, 1, 3])
    np.testing.assert_array_equal(ids, [7, 2, 5])


def test_dataset_mirror_matches_oracle_kg():
    from kelpie_b200.data import Dataset
    z, meta, kg, w, order = load("ComplEx")
    ds = Dataset("g", z["train"], z["valid"], z["test"], kg.num_entities, kg.num_relations)
    assert ds.relation_to_type == kg.relation_to_type
    for e in range(0, kg.num_entities, 17):
        assert set(ds.entity_to_training_triples[e]) == set(kg.facts_of[e])
        assert ds.entity_to_degree.get(e, 0) == kg.degree.get(e, 0)
    np.testing.assert_array_equal(ds.invert_triples(z["test"][:5]), kg.invert(z["test"][:5]))


def test_shard_bounds_balance_and_cover():
    from kelpie_b200.parallel import shard_bounds
    costs = [1, 9, 1, 1, 4, 4, 10, 2]
    b = shard_bounds(costs, 4)
    assert b[0] == 0 and b[-1] == len(costs) and all(x <= y for x, y in zip(b, b[1:]))
    assert shard_bounds([], 3) == [0, 0, 0, 0]
    assert shard_bounds([5], 2)[-1] == 1


WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from kelpie_b200.parallel import ShardedEngine

class FakeEngine:  # MockEngine pattern of the reference's builder tests (test_stochastic_builder.py:7-11)
    """Draws one random number per rule -- for EVERY rule, owned or not, as the real engines do -- so the test also
    checks that the generators of all ranks end where the single-process run leaves them."""
    def compute_relevances(self, pred, rules, snapshots=False, owned=None):
        lo, hi = owned if owned is not None else (0, len(rules))
        noise = [float(torch.rand(1)) for _ in rules]
        rels = [float(sum(t[2] for t in r)) + 0.5 + noise[i] for i, r in enumerate(rules)][lo:hi]
        return (rels, [torch.get_rng_state() for _ in rules]) if snapshots else rels
    def compute_relevance(self, pred, rule):
        return self.compute_relevances(pred, [rule])[0]

dist.init_process_group("gloo")
rules = [[(1, 0, i)] * (1 + i % 3) for i in range(11)]
torch.manual_seed(5)
eng = ShardedEngine(FakeEngine())
got = eng.compute_relevances((1, 0, 2), rules)
tail = float(torch.rand(1))
got2, snaps = eng.compute_relevances((1, 0, 2), rules, snapshots=True)   # the builder's calling convention
assert len(got2) == len(rules) == len(snaps)
assert isinstance(eng.compute_relevance((1, 0, 2), rules[0]), float)
torch.manual_seed(5)
want = FakeEngine().compute_relevances((1, 0, 2), rules)
assert got == want, (got, want)
assert tail == float(torch.rand(1))   # same generator position as the single-process run
dist.barrier()
dist.destroy_process_group()
print("ok", dist.is_initialized())
'''


def test_sharded_engine_two_gloo_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out


def test_full_epoch_draw_matches_reference_order():
    """plans.draw_transe_full_epoch = pairwise_ranking_optimizer.py:100-118: shuffle, randint(2), randint(N) over
    ratio * n samples, first n used."""
    import torch
    from kelpie_b200 import plans
    rng = np.random.default_rng(0)
    rows = np.stack([rng.integers(0, 50, 37), rng.integers(0, 6, 37), rng.integers(0, 50, 37)], 1).astype(np.int64)
    a = rows.copy()
    np.random.seed(4); torch.manual_seed(4)
    pos, neg = plans.draw_transe_full_epoch(a, 50, 5)
    b = rows.copy()
    np.random.seed(4); torch.manual_seed(4)
    np.random.shuffle(b)
    rep = np.repeat(b, 5, axis=0)
    coin = torch.randint(high=2, size=(len(rep),)).numpy()
    rnd = torch.randint(high=50, size=(len(rep),)).numpy()
    want_neg = rep.copy()
    want_neg[coin == 1, 0] = rnd[coin == 1]
    want_neg[coin != 1, 2] = rnd[coin != 1]
    assert np.array_equal(a, b)                       # shuffled in place, state carries to the next epoch
    assert np.array_equal(pos, rep[:37]) and np.array_equal(neg, want_neg[:37])


def test_kelpie_epoch_draws_match_reference_order():
    """plans.draw_transe = pairwise_ranking_optimizer.py:165-195 over all epochs: np.random.shuffle of the ROWS
    (cumulative), randint(N+1) then randint(2) over ratio * n samples, first n of the repeated rows used."""
    import torch
    from kelpie_b200 import plans
    rng = np.random.default_rng(1)
    facts = np.stack([rng.integers(0, 40, 7), rng.integers(0, 3, 7), rng.integers(0, 40, 7)], 1).astype(np.int64)
    hp = dict(epochs=6, negative_triples_ratio=5)
    np.random.seed(9); torch.manual_seed(9)
    n, pos, neg = plans.draw_transe(facts, 3, 41, hp)
    np.random.seed(9); torch.manual_seed(9)
    inv = facts[:, [2, 1, 0]].copy(); inv[:, 1] += 3
    rows = np.vstack((facts, inv))
    assert n == len(rows) == 14
    for e in range(6):
        np.random.shuffle(rows)
        rep = np.repeat(rows, 5, axis=0)
        rnd = torch.randint(high=41, size=(len(rep),)).numpy()
        coin = torch.randint(high=2, size=(len(rep),)).numpy()
        want = rep.copy()
        want[coin == 1, 0] = rnd[coin == 1]
        want[coin != 1, 2] = rnd[coin != 1]
        assert np.array_equal(pos[e * n:(e + 1) * n], rep[:n]) and np.array_equal(neg[e * n:(e + 1) * n], want[:n])
    # opt-in vectorised draws: same shapes / value ranges, every epoch a permutation of the rows
    n2, pos2, neg2 = plans.draw_transe(facts, 3, 41, hp, fast_rng=np.random.default_rng(0))
    assert n2 == n and pos2.shape == pos.shape and neg2.shape == neg.shape
    assert set(map(tuple, pos2)) <= set(map(tuple, np.vstack((facts, inv))))
    changed = (neg2 != pos2)
    assert not changed[:, 1].any() and not (changed[:, 0] & changed[:, 2]).any() and neg2.max() <= 40


def test_native_replay_reproduces_torch_and_numpy_generators():
    """kp_host_rng.cu (kp_mt19937_words / kp_replay_transe_corruptions / kp_replay_numpy_shuffles) against the generators
    themselves: same numbers as the reference's per-epoch calls (pairwise_ranking_optimizer.py:167-172,
    multiclass_nll_optimizer.py:148) and both generators left where those calls leave them."""
    import torch
    from kelpie_b200 import plans
    assert plans.HostReplay.available()
    for seed, E, n, ratio, high in [(0, 65, 4, 5, 24621), (1, 59, 130, 5, 96001), (2, 3, 700, 5, 2), (3, 7, 1, 1, (1 << 28) - 1),
                                    (4, 2, 1300, 5, 1000001)]:
        m = ratio * n
        np.random.seed(seed); torch.manual_seed(seed)
        torch.rand(1, seed + 1)                       # not at a block boundary
        perm = plans.HostReplay.numpy_shuffles(E, n)
        code = plans.HostReplay.transe_corruptions(E, m, n, high)
        plans.HostReplay.torch_skip(3 * (n - 1))
        tails = torch.rand(4), np.random.random(3)
        np.random.seed(seed); torch.manual_seed(seed)
        torch.rand(1, seed + 1)
        idx = np.arange(n)
        for e in range(E):
            np.random.shuffle(idx)
            assert np.array_equal(perm[e], idx)
            rnd = torch.randint(high, (m,)).numpy()[:n]
            coin = torch.randint(2, (m,)).numpy()[:n]
            assert np.array_equal(code[e * n:(e + 1) * n].view(np.uint32), (rnd | (coin << 31)).astype(np.uint32))
        for _ in range(3):
            torch.randperm(n)
        assert torch.equal(tails[0], torch.rand(4)) and np.array_equal(tails[1], np.random.random(3))


def test_replay_self_check_restores_the_host_generators():
    import torch
    from kelpie_b200 import plans
    torch.manual_seed(42); np.random.seed(42)
    torch.rand(3); np.random.random(3)
    t0, n0 = torch.get_rng_state().clone(), np.random.get_state()
    saved, plans.HostReplay._ok = plans.HostReplay._ok, None
    try:
        assert plans.HostReplay.available()
    finally:
        plans.HostReplay._ok = saved if saved is not None else plans.HostReplay._ok
    n1 = np.random.get_state()
    assert torch.equal(t0, torch.get_rng_state()) and np.array_equal(n0[1], n1[1]) and n0[2:] == n1[2:]


def test_native_replay_and_per_call_draws_agree():
    """plans.draw_transe_compact / draw_complex: the native replay and the per-call path give the same tables and leave
    the same generator states, job after job (what a candidate batch does)."""
    import torch
    from kelpie_b200 import plans
    rng = np.random.default_rng(3)
    jobs = [np.stack([rng.integers(0, 40, t), rng.integers(0, 3, t), rng.integers(0, 40, t)], 1) for t in (1, 2, 11, 64, 300)]
    hp = dict(epochs=9, negative_triples_ratio=5, batch_size=512)

    def run(native):
        saved = plans.HostReplay._ok
        plans.HostReplay._ok = native
        try:
            np.random.seed(11); torch.manual_seed(11)
            out = []
            for f in jobs:
                out.append(torch.rand(1, 8).numpy())
                out.extend(np.asarray(x) for x in plans.draw_transe_compact(f, 3, 41, hp))
                out.extend(np.asarray(x) for x in plans.draw_complex(f, 3, hp))   # t = 300: 600 rows > batch size
            out.extend([torch.rand(3).numpy(), np.random.random(3)])
            return out
        finally:
            plans.HostReplay._ok = saved

    a, b = run(True), run(False)  # native: one fused call per TransE job (kp_replay_transe_job)
    assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))
    plans.HostReplay.split_calls = True  # the two separate native calls (shuffles, corruptions); KELPIE_HOST_REPLAY=split
    try:
        c = run(True)
    finally:
        plans.HostReplay.split_calls = False
    assert len(a) == len(c) and all(np.array_equal(x, y) and x.dtype == y.dtype for x, y in zip(a, c))


def test_mt19937_words_match_numpy_bit_generator():
    from kelpie_b200 import runtime
    lib = runtime.load_library()
    bg = np.random.MT19937(2024)
    st = bg.state["state"]
    key, pos = st["key"].astype(np.uint32).copy(), np.array([st["pos"]], np.int32)
    out = np.empty(2000, np.uint32)
    assert lib.kp_mt19937_words(key.ctypes.data, pos.ctypes.data, 2000, out.ctypes.data) == 0
    assert np.array_equal(out, bg.random_raw(2000).astype(np.uint32))
    assert lib.kp_mt19937_words(key.ctypes.data, pos.ctypes.data, 1500, None) == 0    # skip only
    bg.random_raw(1500)
    assert lib.kp_mt19937_words(key.ctypes.data, pos.ctypes.data, 10, out.ctypes.data) == 0
    assert np.array_equal(out[:10], bg.random_raw(10).astype(np.uint32))
    bad = np.array([625], np.int32)
    assert lib.kp_mt19937_words(key.ctypes.data, bad.ctypes.data, 1, out.ctypes.data) < 0
    assert lib.kp_replay_transe_corruptions(key.ctypes.data, pos.ctypes.data, 1, 5, 6, 10, out.ctypes.data) < 0  # used > drawn


def test_dataset_edits_follow_reference_semantics():
    """dataset.py:242-280: add / remove keep train array, per-entity lists, degrees and the DIRECT filter key."""
    from kelpie_b200.data import Dataset
    train = np.array([[0, 0, 1], [0, 0, 2], [1, 1, 2], [3, 0, 1]])
    ds = Dataset("t", train, np.zeros((0, 3), int), np.array([[2, 1, 3]]), 4, 2)
    ds.remove_training_triples([(0, 0, 2), (0, 0, 2)])
    assert len(ds.training_triples) == 3 and (0, 0, 2) not in ds.entity_to_training_triples[0]
    assert ds.to_filter[(0, 0)] == [1] and ds.train_to_filter[(0, 0)] == [1]
    assert ds.entity_to_degree[0] == 1 and ds.entity_to_degree[2] == 1
    ds.add_training_triples([(2, 0, 3)])
    assert tuple(ds.training_triples[-1]) == (2, 0, 3) and ds.to_filter[(2, 0)] == [3]
    assert (2, 0, 3) in ds.entity_to_training_triples[3] and ds.entity_to_degree[3] == 2
    with pytest.raises(ValueError):
        ds.remove_training_triple((1, 0, 0))           # like list.remove in the reference


def test_filter_facts_equal_the_to_filter_dict_after_edits():
    """Dataset.filter_facts() (input of the device CSR builder) carries the same multiset as the dict
    Dataset.to_filter (dataset.py:131-139), including the reference's edit quirks: an added fact only gets its direct
    key, a removed one loses one occurrence of its direct row while a copy in valid / test keeps it filtered."""
    from kelpie_b200 import runtime
    from kelpie_b200.data import Dataset
    rng = np.random.default_rng(3)
    tr = np.stack([rng.integers(0, 40, 300), rng.integers(0, 5, 300), rng.integers(0, 40, 300)], 1)
    te = np.vstack((tr[:20], np.stack([rng.integers(0, 40, 30), rng.integers(0, 5, 30), rng.integers(0, 40, 30)], 1)))
    ds = Dataset("t", tr, tr[20:30], te, 40, 5)

    def csr_of_facts(rows):
        u = np.unique(np.stack((rows[:, 0].astype(np.int64) * 10 + rows[:, 1], rows[:, 2]), 1), axis=0)
        keys, counts = np.unique(u[:, 0], return_counts=True)
        return keys, np.concatenate(([0], np.cumsum(counts))), u[:, 1].astype(np.int32)

    def same():
        a, b = csr_of_facts(ds.filter_facts()), runtime.filter_csr(ds.to_filter, 10)
        return all(np.array_equal(x, y) for x, y in zip(a, b))

    assert same()
    ds.remove_training_triples([tuple(tr[0]), tuple(tr[25]), tuple(tr[100])])  # tr[0] also in test, tr[25] also in valid
    assert same()
    ds.add_training_triples([(1, 2, 3), (39, 4, 0), (1, 2, 3)])
    assert same()
    ds.remove_training_triple((1, 2, 3))
    assert same()


def test_fused_replay_call_rejects_foreign_generator_layouts():
    """kp_replay_transe_job takes torch's generator as the get_rng_state() buffer: anything but the 5056-byte layout it
    knows (or a corrupt position) is refused instead of being walked."""
    import ctypes
    from kelpie_b200 import runtime
    lib = runtime.load_library()
    key = np.zeros(624, dtype=np.uint32)
    pos = np.array([624], dtype=np.int32)
    out = np.zeros(10, dtype=np.int32)
    good = np.zeros(5056, dtype=np.uint8)
    good[8:12].view(np.int32)[0] = 1  # freshly seeded: left = 1, next = 0
    args = (2, 5, 1, 100, out.ctypes.data, out.ctypes.data)
    assert lib.kp_replay_transe_job(key.ctypes.data, pos.ctypes.data, good.ctypes.data, 5056, *args) == 0
    assert lib.kp_replay_transe_job(key.ctypes.data, pos.ctypes.data, good.ctypes.data, 5048, *args) != 0
    bad = good.copy()
    bad[16:24].view(np.uint64)[0] = 9999  # next beyond the 624 words
    bad[8:12].view(np.int32)[0] = 5
    assert lib.kp_replay_transe_job(key.ctypes.data, pos.ctypes.data, bad.ctypes.data, 5056, *args) != 0
    assert lib.kp_replay_transe_job(key.ctypes.data, pos.ctypes.data, good.ctypes.data, 5056, 2, 5, 0, 100, out.ctypes.data, out.ctypes.data) != 0
