"""Host-side construction of post-training batches (the candidate batches the builders emit).

A *job* is one mimic post-training: the mimic's training facts (ids, mimic id = N), its
initial row and the hyper-parameters.  `draw_job` consumes the host random generators in
EXACTLY the order the reference's Kelpie*Optimizer does and returns the
index tables the CUDA kernels read, so that a batch of jobs drawn one after the other
reproduces the reference's sequential run.
"""
import ctypes
import os
from collections import defaultdict

import numpy as np
import torch


def _rows_with_inverses(facts, num_relations):
    """triples + inverse triples (dataset.py:319-331; *_optimizer.py train())."""
    f = np.asarray(facts, dtype=np.int64).reshape(-1, 3)
    t = len(f)
    out = np.empty((2 * t, 3), dtype=np.int64)
    out[:t] = f
    out[t:, 0] = f[:, 2]
    out[t:, 1] = f[:, 1] + num_relations
    out[t:, 2] = f[:, 0]
    return out


class HostReplay:
    """Native replay of the reference's per-epoch generator calls (kp_host_rng.cu behind the C ABI).

    torch's default CPU generator and numpy's legacy global RandomState are both 32-bit Mersenne Twisters and
    `torch.randint(high < 2^32)`, `torch.randperm` and `np.random.shuffle` are fixed functions of consecutive output
    words, so the 3 * epochs Python calls of one TransE post-training (2 * epochs randint, epochs shuffles) become two
    native calls that produce the SAME numbers and leave both generators advanced as the reference's calls would (the
    next `torch.rand` of an init row continues from there).  torch's state travels through `get_rng_state` /
    `set_rng_state` (seed u64, left i32, seeded i32, next u64, 624 key words stored as u64, then the normal-distribution
    cache: 5056 bytes); numpy's is walked in place through the bit generator's `ctypes.state_address` (624 key words,
    then the position).  `available()` checks every replay against torch / numpy themselves once per process, on the
    live and on a freshly seeded state; where one does not reproduce (another torch or numpy, another generator) the
    callers keep the per-call path."""

    _N, _SIZE = 624, 5056
    split_calls = os.environ.get("KELPIE_HOST_REPLAY") == "split"  # A/B: two native calls per TransE job instead of the fused one
    MAX_HIGH = 1 << 28  # torch 2.11 reduces 64-bit words (two generator words per element) from this range on
    _ok = None
    _lib = None
    _np_addr = None
    _np_bg = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            from . import runtime
            cls._lib = runtime.load_library()
        return cls._lib

    # ---- torch's default CPU generator
    @classmethod
    def torch_state(cls):
        buf = torch.get_rng_state().numpy()
        if buf.size != cls._SIZE:
            raise RuntimeError("unknown torch CPU generator state layout")
        key = buf[24:24 + 8 * cls._N].view(np.uint64).astype(np.uint32)
        left, nxt = int(buf[8:12].view(np.int32)[0]), int(buf[16:24].view(np.uint64)[0])
        pos = np.array([cls._N if (left == 1 and nxt == 0) else nxt], dtype=np.int32)  # freshly seeded: regenerate first
        return buf, key, pos

    @classmethod
    def torch_commit(cls, buf, key, pos):
        buf[24:24 + 8 * cls._N].view(np.uint64)[:] = key
        buf[16:24].view(np.uint64)[0] = int(pos[0])
        buf[8:12].view(np.int32)[0] = cls._N + 1 - int(pos[0])  # torch keeps left + next == 625
        buf[12:16].view(np.int32)[0] = 1
        torch.set_rng_state(torch.from_numpy(buf))

    @classmethod
    def torch_skip(cls, count):
        """Advance torch's generator by `count` words (E calls of torch.randperm(n): E * (n - 1))."""
        buf, key, pos = cls.torch_state()
        rc = cls.lib().kp_mt19937_words(key.ctypes.data, pos.ctypes.data, int(count), None)
        assert rc == 0, rc
        cls.torch_commit(buf, key, pos)

    @classmethod
    def transe_corruptions(cls, E, drawn, used, high):
        """neg_code [E * used] int32 of `for e: torch.randint(high, (drawn,)); torch.randint(2, (drawn,))`, first `used`."""
        code = np.empty(E * used, dtype=np.int32)
        buf, key, pos = cls.torch_state()
        rc = cls.lib().kp_replay_transe_corruptions(key.ctypes.data, pos.ctypes.data, E, drawn, used, high, code.ctypes.data)
        assert rc == 0, rc
        cls.torch_commit(buf, key, pos)
        return code

    @classmethod
    def transe_job(cls, E, n, ratio, high):
        """numpy_shuffles + transe_corruptions of one job in ONE native call (kp_replay_transe_job), torch's generator state
        handed over as the get_rng_state() buffer itself: (pos_idx [E * n] int32, neg_code [E * n] int32)."""
        pos_idx = np.empty(E * n, dtype=np.int32)
        code = np.empty(E * n, dtype=np.int32)
        state = torch.get_rng_state()
        addr = cls._numpy_addr()
        rc = cls.lib().kp_replay_transe_job(addr, addr + 4 * cls._N, state.data_ptr(), state.numel(), E, n, ratio, high,
                                            pos_idx.ctypes.data, code.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"kp_replay_transe_job failed ({rc}): unknown torch CPU generator state layout?")
        torch.set_rng_state(state)
        return pos_idx, code

    # ---- numpy's legacy global RandomState
    @classmethod
    def _numpy_addr(cls):
        """Address of the global RandomState's generator words (624 x u32, then the position); re-resolved if the
        bit generator object was swapped (np.random.set_bit_generator)."""
        bg = np.random.mtrand._rand._bit_generator
        if bg is not cls._np_bg:
            if type(bg).__name__ != "MT19937":
                raise RuntimeError("numpy's global generator is not an MT19937")
            addr = bg.ctypes.state_address
            cls._np_bg, cls._np_addr = bg, int(getattr(addr, "value", addr))  # the reference keeps the state alive
            if cls._ok:  # a generator object the replay has not been checked against yet
                cls._ok = None
        return cls._np_addr

    @classmethod
    def numpy_shuffles(cls, E, n):
        """[E, n] int32: an index vector after each of E cumulative np.random.shuffle calls."""
        addr = cls._numpy_addr()
        perm = np.empty((E, n), dtype=np.int32)
        rc = cls.lib().kp_replay_numpy_shuffles(addr, addr + 4 * cls._N, E, n, perm.ctypes.data)
        assert rc == 0, rc
        return perm

    @classmethod
    def numpy_snapshot(cls):
        """The global RandomState's generator words + position as bytes (a 2.5 KB copy instead of the 80 us
        np.random.get_state() tuple); the legacy Gaussian cache is not part of it -- nothing on this path draws normals
        from numpy between a snapshot and its restore."""
        return ctypes.string_at(cls._numpy_addr(), 4 * cls._N + 4)

    @classmethod
    def numpy_restore(cls, raw):
        ctypes.memmove(cls._numpy_addr(), raw, 4 * cls._N + 4)

    @classmethod
    def available(cls):
        if cls._ok and np.random.mtrand._rand._bit_generator is not cls._np_bg:
            cls._ok = None  # numpy's global generator object was swapped since the check
        if cls._ok is None and os.environ.get("KELPIE_HOST_REPLAY") == "0":  # A/B switch: per-call draws
            cls._ok = False
        if cls._ok is None:
            saved_t, saved_n = torch.get_rng_state(), np.random.get_state()
            try:
                ok = True
                for seed in (None, 12345):  # the live states, and freshly seeded ones
                    if seed is not None:
                        torch.default_generator.manual_seed(seed)  # the CPU generator only: torch.manual_seed reseeds CUDA's too
                        np.random.seed(seed)
                    start_t, start_n = torch.get_rng_state(), np.random.get_state()
                    code = cls.transe_corruptions(3, 700, 650, 24621)
                    cls.torch_skip(2 * 9)
                    tail_t = torch.rand(5)
                    snap = cls.numpy_snapshot()
                    perm = cls.numpy_shuffles(3, 37)
                    tail_n = np.random.random(3)
                    cls.numpy_restore(snap)
                    ok = ok and np.array_equal(perm, cls.numpy_shuffles(3, 37)) and np.array_equal(tail_n, np.random.random(3))
                    torch.set_rng_state(start_t)
                    np.random.set_state(start_n)
                    for e in range(3):
                        rnd = torch.randint(24621, (700,)).numpy()[:650]
                        coin = torch.randint(2, (700,)).numpy()[:650]
                        ref = (rnd | (coin << 31)).astype(np.uint32).view(np.int32)
                        ok = ok and np.array_equal(code[e * 650:(e + 1) * 650], ref)
                    torch.randperm(10), torch.randperm(10)
                    ok = ok and torch.equal(tail_t, torch.rand(5))
                    idx = np.arange(37)
                    for e in range(3):
                        np.random.shuffle(idx)
                        ok = ok and np.array_equal(perm[e], idx)
                    ok = ok and np.array_equal(tail_n, np.random.random(3))
                    # the fused per-job call against torch / numpy themselves, and where it leaves both generators
                    torch.set_rng_state(start_t)
                    np.random.set_state(start_n)
                    pi, cd = cls.transe_job(2, 11, 5, 24621)
                    end_t, end_n = torch.rand(3), np.random.random(2)
                    torch.set_rng_state(start_t)
                    np.random.set_state(start_n)
                    idx = np.arange(11)
                    for e in range(2):
                        np.random.shuffle(idx)
                        rnd = torch.randint(24621, (55,)).numpy()[:11]
                        coin = torch.randint(2, (55,)).numpy()[:11]
                        ref = (rnd | (coin << 31)).astype(np.uint32).view(np.int32)
                        ok = ok and np.array_equal(pi[e * 11:(e + 1) * 11], idx[np.arange(11) // 5])
                        ok = ok and np.array_equal(cd[e * 11:(e + 1) * 11], ref)
                    ok = ok and torch.equal(end_t, torch.rand(3)) and np.array_equal(end_n, np.random.random(2))
                cls._ok = bool(ok)
            except Exception:
                cls._ok = False
            finally:
                torch.set_rng_state(saved_t)
                np.random.set_state(saved_n)
        return cls._ok


def _randint_epochs(E, m, n, high):
    """`for e in range(E): a = torch.randint(high, (m,)); b = torch.randint(2, (m,))`, the first n of each, call by call
    (the path kept for a torch whose generator HostReplay does not reproduce; one large draw with a single range
    cannot reproduce them)."""
    rnd_t = torch.empty((E, m), dtype=torch.int64)
    coin_t = torch.empty((E, m), dtype=torch.int64)
    for e in range(E):
        torch.randint(high, (m,), out=rnd_t[e])
        torch.randint(2, (m,), out=coin_t[e])
    return rnd_t.numpy()[:, :n], coin_t.numpy()[:, :n]


def draw_transe_compact(facts, num_relations, n_ent_with_mimic, hp, fast_rng=None):
    """pairwise_ranking_optimizer.py:165-195 as COMPACT index tables (kelpie_b200.h, kp_pt_batch.pos_idx / neg_code):
    returns (rows_per_epoch n, rows [n,3] int32 = triples + inverses, pos_idx [E*n] = which row is the positive,
    neg_code [E*n] int32 = corrupting entity | head-corrupted << 31).

    Reference order (default): per epoch np.random.shuffle(rows) (cumulative: an index vector is shuffled, which
    draws the same numbers and applies the same permutation as shuffling the rows), torch.randint(N+1) then
    torch.randint(2) over ratio * n samples of which the first n are used.  The per-epoch Python work is kept to
    those three generator calls; the tables are assembled once for all epochs.
    fast_rng (opt-in, a numpy Generator): independent uniform permutations / corruptions for all epochs in three
    vectorised draws -- the same distribution, not the reference's random numbers."""
    rows = _rows_with_inverses(facts, num_relations)
    n, E, ratio = len(rows), int(hp["epochs"]), int(hp["negative_triples_ratio"])
    if n == 0:
        return 0, np.zeros((0, 3), np.int32), np.zeros(0, np.int64), np.zeros(0, np.int32)
    take = np.arange(n) // ratio  # first n rows of np.repeat(rows, ratio)
    if fast_rng is not None:
        perm = np.argsort(fast_rng.random((E, n)), axis=1)
        rnd = fast_rng.integers(0, n_ent_with_mimic, (E, n))
        coin = fast_rng.integers(0, 2, (E, n))
    elif n_ent_with_mimic < HostReplay.MAX_HIGH and HostReplay.available():
        # numpy's generator is independent of torch's: interleaving the shuffles with the corruptions draws the same numbers
        if not HostReplay.split_calls:
            pos_idx, code = HostReplay.transe_job(E, n, ratio, n_ent_with_mimic)
            return n, rows.astype(np.int32), pos_idx, code
        perm = HostReplay.numpy_shuffles(E, n)  # A/B: the two separate native calls
        code = HostReplay.transe_corruptions(E, ratio * n, n, n_ent_with_mimic)
        return n, rows.astype(np.int32), perm[:, take].reshape(-1), code
    else:
        perm = np.empty((E, n), dtype=np.int64)
        idx = np.arange(n)
        for e in range(E):
            np.random.shuffle(idx)
            perm[e] = idx
        rnd, coin = _randint_epochs(E, ratio * n, n, n_ent_with_mimic)
    code = (rnd.astype(np.int64) | ((coin == 1).astype(np.int64) << 31)).astype(np.uint32).view(np.int32)
    return n, rows.astype(np.int32), perm[:, take].reshape(-1), code.reshape(-1)


def expand_transe(rows, pos_idx, neg_code):
    """Compact tables -> the full (pos, neg) [.,3] int32 row tables the reference's step_on_batch sees."""
    pos = rows[pos_idx].astype(np.int32).reshape(-1, 3)
    neg = pos.copy()
    code = np.asarray(neg_code).view(np.uint32)
    head, rnd = (code >> 31) == 1, (code & 0x7FFFFFFF).astype(np.int32)
    neg[:, 0] = np.where(head, rnd, pos[:, 0])
    neg[:, 2] = np.where(head, pos[:, 2], rnd)
    return pos, neg


def draw_transe(facts, num_relations, n_ent_with_mimic, hp, fast_rng=None):
    """As draw_transe_compact, expanded: returns (rows_per_epoch, pos [E*n,3], neg [E*n,3])."""
    n, rows, pos_idx, code = draw_transe_compact(facts, num_relations, n_ent_with_mimic, hp, fast_rng)
    if n == 0:
        return 0, np.zeros((0, 3), np.int32), np.zeros((0, 3), np.int32)
    pos, neg = expand_transe(rows, pos_idx, code)
    return n, pos, neg


def draw_transe_full_epoch(rows, num_entities, ratio):
    """One epoch of the FULL-model trainer (pairwise_ranking_optimizer.py:100-118): shuffles `rows` in place
    (np.random, the state carries across epochs), then draws head_or_tail = randint(2) and the corrupting
    entities = randint(num_entities) over ratio * n samples, of which -- like the training loop (:127-134) --
    only the first n are used.  Returns (pos [n,3], neg [n,3]) int32."""
    n = len(rows)
    np.random.shuffle(rows)
    coin = torch.randint(high=2, size=(ratio * n,))[:n].numpy()
    rnd = torch.randint(high=num_entities, size=(ratio * n,))[:n].numpy()
    pos = rows[np.arange(n) // ratio].astype(np.int32)
    neg = pos.copy()
    head = coin == 1
    neg[head, 0] = rnd[head]
    neg[~head, 2] = rnd[~head]
    return pos, neg


def draw_complex(facts, num_relations, hp, fast_rng=None):
    """multiclass_nll_optimizer.py:147-164.  Returns (rows_per_epoch, rows, static_epochs)."""
    rows = _rows_with_inverses(facts, num_relations)
    n, E = len(rows), int(hp["epochs"])
    if fast_rng is not None:
        if n <= int(hp["batch_size"]):
            return n, rows.astype(np.int32), True
        perm = np.argsort(fast_rng.random((E, n)), axis=1)
        return n, rows[perm].reshape(-1, 3).astype(np.int32), False
    if n <= int(hp["batch_size"]):
        # one step per epoch over ALL rows: the permutation cannot change a mean over the
        # batch, so one epoch of rows is reused; the generator is still advanced as the
        # reference's torch.randperm would.
        if HostReplay.available():  # torch.randperm(n) consumes n - 1 words of the generator
            if n > 1:
                HostReplay.torch_skip(E * (n - 1))
        else:
            for _ in range(E):
                torch.randperm(n)
        return n, rows.astype(np.int32), True
    out = np.empty((E, n, 3), dtype=np.int32)
    for e in range(E):
        out[e] = rows[torch.randperm(n).numpy()]
    return n, out.reshape(-1, 3), False


def plan_conve(facts, num_relations):
    """bce_optimizer.py:92-96: (lhs, rel) pairs in first-seen order and their positives."""
    rows = _rows_with_inverses(facts, num_relations)
    vocab = defaultdict(list)
    for s, p, o in rows:
        vocab[(int(s), int(p))].append(int(o))
    pairs = np.array([(s, p, 0) for s, p in vocab], dtype=np.int32).reshape(-1, 3)
    lens = [len(set(v)) for v in vocab.values()]
    ids = np.array([x for v in vocab.values() for x in sorted(set(v))], dtype=np.int32)
    return pairs, np.array(lens, dtype=np.int64), ids


def _host_buffer(shape, dtype):
    """Flat batch arrays are assembled straight into page-locked memory when a CUDA device is present, so the
    host-to-device copy is one DMA without a staging pass (runtime.Context.dev recognises pinned arrays)."""
    if torch.cuda.is_available():
        t = torch.empty(tuple(shape), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
        return t.numpy()
    return np.empty(shape, dtype=dtype)


def _concat(parts, dtype):
    parts = [np.asarray(p) for p in parts]
    n = sum(len(p) for p in parts)
    out = _host_buffer((n,) + parts[0].shape[1:], dtype)
    np.concatenate(parts, out=out, casting="unsafe")
    return out


class Batch:
    """Accumulates jobs and turns them into the flat arrays of kp_pt_batch."""

    def __init__(self, kind, num_entities, num_relations, hp, fast_rng=None):
        self.kind, self.N, self.R, self.hp = kind, int(num_entities), int(num_relations), hp
        self.fast_rng = fast_rng  # None: the reference's generators in the reference's order
        self.init_rows, self.rows_per_epoch = [], []
        self.pos, self.neg, self.pos_lens, self.pos_ids = [], [], [], []
        self.facts = []  # TransE: each job's distinct rows (triples + inverses)
        self.statics = []

    def __len__(self):
        return len(self.init_rows)

    def add(self, facts, init_row):
        """Draw (in reference order) and append one job; returns its index in the batch."""
        if self.kind == "TransE":  # kept compact; arrays() expands on request
            n, rows, pos_idx, code = draw_transe_compact(facts, self.R, self.N + 1, self.hp, self.fast_rng)
            self.facts.append(rows)
            self.pos.append(pos_idx)
            self.neg.append(code)
            static = False
        elif self.kind == "ComplEx":
            n, pos, static = draw_complex(facts, self.R, self.hp, self.fast_rng)
            self.pos.append(pos)
        else:
            pos, lens, ids = plan_conve(facts, self.R)
            n, static = len(pos), True
            self.pos.append(pos)
            self.pos_lens.append(lens)
            self.pos_ids.append(ids)
        self.statics.append(static)
        self.rows_per_epoch.append(n)
        if isinstance(init_row, torch.Tensor) and init_row.is_cuda:  # drawn on the device (TransE's xavier_normal_): stays there
            self.init_rows.append(init_row.detach().to(torch.float32).reshape(-1))
        else:
            self.init_rows.append(np.asarray(init_row, dtype=np.float32).reshape(-1))
        return len(self.init_rows) - 1

    def _stacked_init_rows(self):
        """[C, D] fp32: numpy when every row came from the host, one device tensor when any was drawn on the device."""
        dev = [r for r in self.init_rows if isinstance(r, torch.Tensor)]
        if not dev:
            return np.stack(self.init_rows).astype(np.float32)
        d = dev[0].device
        return torch.stack([r if isinstance(r, torch.Tensor) else torch.from_numpy(r).to(d) for r in self.init_rows])

    def arrays(self, compact=False):
        """Flat arrays of kp_pt_batch (keyword arguments of runtime.Context.post_train).  compact (TransE only): the
        6-bytes-per-row index tables (fact_off / facts / pos_idx / neg_code) instead of pos / neg."""
        if self.kind == "TransE":
            return self._arrays_transe(compact and all(len(f) < 65536 for f in self.facts))
        static = all(self.statics)
        if not static:  # mixed batch: unroll the single-epoch jobs so every job is epoch-major
            E = int(self.hp["epochs"])
            self.pos = [np.tile(p, (E, 1)) if st else p for p, st in zip(self.pos, self.statics)]
            self.statics = [False] * len(self.statics)
        sizes = [len(p) for p in self.pos]
        row_off = np.zeros(len(sizes) + 1, dtype=np.int64)
        row_off[1:] = np.cumsum(sizes)
        out = dict(
            init_rows=self._stacked_init_rows(),
            row_off=row_off,
            rows_per_epoch=np.array(self.rows_per_epoch, dtype=np.int32),
            pos=_concat(self.pos, np.int32) if sum(sizes) else np.zeros((1, 3), np.int32),
            static_epochs=static,
        )
        if self.kind == "ConvE":
            lens = np.concatenate(self.pos_lens) if self.pos_lens else np.zeros(0, np.int64)
            off = np.zeros(len(lens) + 1, dtype=np.int64)
            off[1:] = np.cumsum(lens)
            out["pos_off"] = off
            out["pos_ids"] = np.concatenate(self.pos_ids) if len(lens) else np.zeros(1, np.int32)
        return out

    def _arrays_transe(self, compact):
        sizes = [len(p) for p in self.pos]
        row_off = np.zeros(len(sizes) + 1, dtype=np.int64)
        row_off[1:] = np.cumsum(sizes)
        out = dict(
            init_rows=self._stacked_init_rows(),
            row_off=row_off,
            rows_per_epoch=np.array(self.rows_per_epoch, dtype=np.int32),
            static_epochs=False,
        )
        total = int(row_off[-1])
        if compact:
            fact_off = np.zeros(len(sizes) + 1, dtype=np.int64)
            fact_off[1:] = np.cumsum([len(f) for f in self.facts])
            out["pos"] = None
            out["fact_off"] = fact_off
            out["facts"] = _concat(self.facts, np.int32) if fact_off[-1] else np.zeros((1, 3), np.int32)
            out["pos_idx"] = _concat(self.pos, np.uint16) if total else np.zeros(1, np.uint16)
            out["neg_code"] = _concat(self.neg, np.int32) if total else np.zeros(1, np.int32)
        elif total:
            full = [expand_transe(f, p, c) for f, p, c in zip(self.facts, self.pos, self.neg) if len(p)]
            out["pos"] = _concat([x[0] for x in full], np.int32)
            out["neg"] = _concat([x[1] for x in full], np.int32)
        else:
            out["pos"] = out["neg"] = np.zeros((1, 3), np.int32)
        return out
