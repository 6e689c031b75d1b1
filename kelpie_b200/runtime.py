"""ctypes binding of libkelpie_b200.so (include/kelpie_b200.h) for PyTorch callers.

PyTorch is plumbing here: it owns device memory and streams; every computation on the hot
path happens inside the library's CUDA kernels.  There is no CPU fallback: if the library
is missing or no sm_100 device is present, construction raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

import numpy as np
import torch

KP_TRANSE, KP_COMPLEX, KP_CONVE = 0, 1, 2
KIND_ID = {"TransE": KP_TRANSE, "ComplEx": KP_COMPLEX, "ConvE": KP_CONVE}
RANK_ENGINE_MIN, RANK_ENGINE_MAX, RANK_MODEL, RANK_CONVE_SORT = 0, 1, 2, 3
OPT_ADAGRAD, OPT_ADAM, OPT_SGD = 0, 1, 2

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libkelpie_b200.so")

EXPORTS = [
    "kp_ctx_create", "kp_ctx_destroy", "kp_last_error", "kp_abi_version", "kp_filter_upload",
    "kp_filter_build", "kp_filter_download",
    "kp_all_scores", "kp_score_triples", "kp_filtered_rank", "kp_post_train_batch", "kp_launch_count", "kp_set_option", "kp_stat",
    "kp_debug_contract", "kp_dp_relevance",
    "kp_mt19937_words", "kp_replay_transe_corruptions", "kp_replay_numpy_shuffles", "kp_replay_transe_job",
    "kp_transe_fit_create", "kp_transe_fit_steps", "kp_transe_fit_destroy", "kp_transe_fit_error", "kp_transe_fit_launches",
    "kp_complex_fit_create", "kp_complex_fit_steps", "kp_complex_fit_destroy", "kp_complex_fit_error", "kp_complex_fit_launches",
    "kp_conve_fit_create", "kp_conve_fit_steps", "kp_conve_fit_destroy", "kp_conve_fit_error", "kp_conve_fit_launches",
]


class ConvEWeights(Structure):
    _fields_ = [
        ("conv_w", c_void_p), ("conv_b", c_void_p), ("fc_w", c_void_p), ("fc_b", c_void_p),
        ("bn1", c_void_p), ("bn2", c_void_p), ("bn3", c_void_p),
        ("n_filters", c_int32), ("hidden", c_int32),
        ("drop_input", c_float), ("drop_feature", c_float), ("drop_hidden", c_float),
    ]


class HP(Structure):
    _fields_ = [
        ("epochs", c_int32), ("batch_size", c_int32), ("optimizer", c_int32),
        ("lr", c_float), ("beta1", c_float), ("beta2", c_float), ("eps", c_float),
        ("margin", c_float), ("reg_weight", c_float), ("label_smoothing", c_float),
        ("regularizer", c_int32),  # ABI 3: 0 = N3, 1 = N2 (ComplEx)
    ]


class PTBatch(Structure):
    _fields_ = [
        ("n_candidates", c_int32), ("static_epochs", c_int32),
        ("max_rows_per_epoch", c_int32), ("reserved", c_int32), ("total_rows", c_int64),
        ("row_off", c_void_p), ("rows_per_epoch", c_void_p), ("pos", c_void_p), ("neg", c_void_p),
        ("pos_off", c_void_p), ("pos_ids", c_void_p), ("init_rows", c_void_p), ("out_rows", c_void_p),
        ("dropout_seed", c_uint64),
        ("fact_off", c_void_p), ("facts", c_void_p), ("pos_idx", c_void_p), ("neg_code", c_void_p),  # TransE compact tables
    ]


_lib = None


def load_library():
    """Load the in-tree shared library; raise (never fall back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            f"{_LIB_PATH} is missing: build it with `python -m kelpie_b200.build` "
            "(kelpie_b200 has no CPU or PyTorch fallback path)"
        )
    lib = ctypes.CDLL(_LIB_PATH)
    lib.kp_ctx_create.argtypes = [c_int, c_int, c_int64, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                                  POINTER(ConvEWeights), POINTER(c_void_p)]
    lib.kp_ctx_create.restype = c_int
    lib.kp_ctx_destroy.argtypes = [c_void_p]
    lib.kp_ctx_destroy.restype = c_int
    lib.kp_last_error.argtypes = [c_void_p]
    lib.kp_last_error.restype = c_char_p
    lib.kp_abi_version.restype = c_int
    lib.kp_filter_upload.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]
    lib.kp_filter_upload.restype = c_int
    lib.kp_filter_build.argtypes = [c_void_p, c_int64, c_void_p, c_void_p]
    lib.kp_filter_build.restype = c_int
    lib.kp_filter_download.argtypes = [c_void_p, POINTER(c_int64), POINTER(c_int64), c_void_p, c_void_p, c_void_p]
    lib.kp_filter_download.restype = c_int
    lib.kp_all_scores.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]
    lib.kp_all_scores.restype = c_int
    lib.kp_score_triples.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.kp_score_triples.restype = c_int
    lib.kp_filtered_rank.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.kp_filtered_rank.restype = c_int
    lib.kp_post_train_batch.argtypes = [c_void_p, POINTER(PTBatch), POINTER(HP), c_void_p]
    lib.kp_post_train_batch.restype = c_int
    lib.kp_launch_count.argtypes = [c_void_p]
    lib.kp_launch_count.restype = c_int64
    lib.kp_set_option.argtypes = [c_void_p, c_char_p, c_int64]
    lib.kp_set_option.restype = c_int
    lib.kp_stat.argtypes = [c_void_p, c_char_p, POINTER(ctypes.c_double)]
    lib.kp_stat.restype = c_int
    lib.kp_debug_contract.argtypes = [c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.kp_debug_contract.restype = c_int
    lib.kp_dp_relevance.argtypes = [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, ctypes.c_float, ctypes.c_float, c_int32,
                                    c_void_p, c_void_p]
    lib.kp_dp_relevance.restype = c_int
    lib.kp_mt19937_words.argtypes = [c_void_p, c_void_p, c_int64, c_void_p]
    lib.kp_mt19937_words.restype = c_int
    lib.kp_replay_transe_corruptions.argtypes = [c_void_p, c_void_p, c_int32, c_int64, c_int64, ctypes.c_uint32, c_void_p]
    lib.kp_replay_transe_corruptions.restype = c_int
    lib.kp_replay_numpy_shuffles.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_void_p]
    lib.kp_replay_numpy_shuffles.restype = c_int
    lib.kp_replay_transe_job.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, ctypes.c_uint32, c_void_p, c_void_p]
    lib.kp_replay_transe_job.restype = c_int
    lib.kp_transe_fit_create.argtypes = [c_int, c_int64, c_int64, c_int32, c_int32, ctypes.c_float, ctypes.c_float,
                                         ctypes.c_float, c_void_p, c_void_p, POINTER(c_void_p)]
    lib.kp_transe_fit_create.restype = c_int
    lib.kp_transe_fit_steps.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.kp_transe_fit_steps.restype = c_int
    lib.kp_transe_fit_destroy.argtypes = [c_void_p]
    lib.kp_transe_fit_destroy.restype = c_int
    lib.kp_transe_fit_error.argtypes = [c_void_p]
    lib.kp_transe_fit_error.restype = c_char_p
    lib.kp_transe_fit_launches.argtypes = [c_void_p]
    lib.kp_transe_fit_launches.restype = c_int64
    lib.kp_complex_fit_create.argtypes = [c_int, c_int64, c_int64, c_int32, c_int32, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                          ctypes.c_float, c_int32, c_void_p, c_void_p, POINTER(c_void_p)]
    lib.kp_complex_fit_create.restype = c_int
    lib.kp_complex_fit_steps.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.kp_complex_fit_steps.restype = c_int
    lib.kp_complex_fit_destroy.argtypes = [c_void_p]
    lib.kp_complex_fit_destroy.restype = c_int
    lib.kp_complex_fit_error.argtypes = [c_void_p]
    lib.kp_complex_fit_error.restype = c_char_p
    lib.kp_complex_fit_launches.argtypes = [c_void_p]
    lib.kp_complex_fit_launches.restype = c_int64
    lib.kp_conve_fit_create.argtypes = [c_int, c_int64, c_int64, c_int32, POINTER(ConvEParams), ctypes.c_float, c_int32, c_int64,
                                        c_void_p, c_void_p, c_void_p, ctypes.c_uint64, POINTER(c_void_p)]
    lib.kp_conve_fit_create.restype = c_int
    lib.kp_conve_fit_steps.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, ctypes.c_float, c_void_p, c_void_p]
    lib.kp_conve_fit_steps.restype = c_int
    lib.kp_conve_fit_destroy.argtypes = [c_void_p]
    lib.kp_conve_fit_destroy.restype = c_int
    lib.kp_conve_fit_error.argtypes = [c_void_p]
    lib.kp_conve_fit_error.restype = c_char_p
    lib.kp_conve_fit_launches.argtypes = [c_void_p]
    lib.kp_conve_fit_launches.restype = c_int64
    _lib = lib
    return lib


def _ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def _np_ptr(a):
    return c_void_p(a.ctypes.data)


def filter_csr(to_filter, n_relations2):
    """Python dict {(entity, relation): [ids...]} -> (keys, offsets, ids) of kp_filter_upload.

    The multiset lists of dataset.py:136-139 are de-duplicated and sorted: masking an id
    twice is the same as masking it once.
    """
    items = [(int(e) * n_relations2 + int(r), sorted(set(int(x) for x in v))) for (e, r), v in to_filter.items() if len(v)]
    items.sort(key=lambda kv: kv[0])
    keys = np.array([k for k, _ in items], dtype=np.int64)
    off = np.zeros(len(items) + 1, dtype=np.int64)
    if items:
        off[1:] = np.cumsum([len(v) for _, v in items])
    ids = np.fromiter((x for _, v in items for x in v), dtype=np.int32, count=int(off[-1]))
    return keys, off, ids


class Context:
    """One kp_ctx: device-resident tables + filter CSR of one model on one GPU."""

    def __init__(self, kind, ent, rel, norm=2, conve=None, device=None):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("kelpie_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.kind = kind
        # keep the tables alive: device tensors are borrowed by the library, not copied
        self.ent = torch.as_tensor(ent, dtype=torch.float32).to(self.device).contiguous()
        self.rel = torch.as_tensor(rel, dtype=torch.float32).to(self.device).contiguous()
        self.N, self.D = self.ent.shape
        self.R2 = self.rel.shape[0]
        cw, self._conve_keep = None, None
        if kind == "ConvE":
            c = {k: torch.as_tensor(v, dtype=torch.float32).contiguous().cpu() for k, v in conve.items() if k != "dropout"}
            bn = lambda i: torch.cat([c[f"bn{i}_w"].view(-1), c[f"bn{i}_b"].view(-1), c[f"bn{i}_mean"].view(-1), c[f"bn{i}_var"].view(-1)]).contiguous()
            keep = dict(conv_w=c["conv_w"], conv_b=c["conv_b"], fc_w=c["fc_w"], fc_b=c["fc_b"], bn1=bn(1), bn2=bn(2), bn3=bn(3))
            drop = conve.get("dropout", (0.0, 0.0, 0.0))
            cw = ConvEWeights(*[c_void_p(keep[k].data_ptr()) for k in ("conv_w", "conv_b", "fc_w", "fc_b", "bn1", "bn2", "bn3")],
                              int(c["conv_w"].shape[0]), int(c["fc_w"].shape[1]), float(drop[0]), float(drop[1]), float(drop[2]))
            self._conve_keep = keep
            self.dropout = tuple(float(x) for x in drop)
        handle = c_void_p()
        rc = self.lib.kp_ctx_create(self.device.index, KIND_ID[kind], self.N, self.R2, self.D, int(norm),
                                    _ptr(self.ent), _ptr(self.rel), ctypes.byref(cw) if cw is not None else None,
                                    ctypes.byref(handle))
        if rc != 0:
            raise RuntimeError(f"kp_ctx_create failed ({rc}): {self.lib.kp_last_error(None).decode()}")
        self.handle = handle
        self.has_filter = False
        # A/B knobs for measurements: KELPIE_B200_OPTS="umma_min_rows=32,cx_merge=0" (kp_set_option names)
        for kv in filter(None, os.environ.get("KELPIE_B200_OPTS", "").split(",")):
            name, value = kv.split("=")
            self.set_option(name.strip(), int(value))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.kp_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.lib.kp_last_error(self.handle).decode()}")

    def _stream(self):
        return c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def dev(self, a, dtype):
        """numpy / tensor -> device tensor of `dtype` (pinned staging for host arrays)."""
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype, non_blocking=True).contiguous()
        t = torch.from_numpy(np.ascontiguousarray(a))
        if t.dtype != dtype:
            t = t.to(dtype)
        return t.pin_memory().to(self.device, non_blocking=True)

    _NP = {torch.float32: np.float32, torch.int32: np.int32, torch.int64: np.int64, torch.uint16: np.uint16}
    PACK_LIMIT = int(os.environ.get("KP_PACK_LIMIT", 1 << 20))  # small calls (explain-path batches) travel as ONE pinned block and ONE H2D copy

    def dev_many(self, items):
        """[(array | tensor | None, torch dtype)] -> device tensors.  Host arrays of a small call are packed into one
        pinned staging block (two alternate, reused) and cross PCIe in a single copy instead of one pin + copy each;
        large or already-pinned arrays and device tensors take the plain path."""
        out, host, total = [None] * len(items), [], 0
        for i, (a, dt) in enumerate(items):
            if a is None:
                continue
            if isinstance(a, torch.Tensor):
                out[i] = a.to(device=self.device, dtype=dt, non_blocking=True).contiguous()
                continue
            arr = np.ascontiguousarray(a, dtype=self._NP[dt])
            host.append((i, arr, dt, total))
            total += (arr.nbytes + 255) & ~255
        if not host:
            return out
        if total > self.PACK_LIMIT:
            for i, arr, dt, _ in host:
                out[i] = self.dev(arr, dt)
            return out
        st = getattr(self, "_stage", None)
        if st is None:
            st = self._stage = {"buf": [None, None], "ev": [None, None], "n": 0}
        k = st["n"] & 1
        st["n"] += 1
        if st["buf"][k] is None or st["buf"][k].numel() < total:
            st["buf"][k] = torch.empty(max(total, 1 << 16), dtype=torch.uint8, pin_memory=True)
            st["ev"][k] = torch.cuda.Event()
        else:
            st["ev"][k].synchronize()  # the copy that last read this block has finished (two calls ago)
        hb = st["buf"][k].numpy()
        for _, arr, _, off in host:
            hb[off:off + arr.nbytes] = arr.reshape(-1).view(np.uint8)
        blk = torch.empty(total, dtype=torch.uint8, device=self.device)
        blk.copy_(st["buf"][k][:total], non_blocking=True)
        st["ev"][k].record()
        for i, arr, dt, off in host:
            out[i] = blk[off:off + arr.nbytes].view(dt).view(arr.shape)
        return out

    @property
    def launches(self):
        return int(self.lib.kp_launch_count(self.handle))

    def set_option(self, name, value):
        self._check(self.lib.kp_set_option(self.handle, name.encode(), int(value)), "kp_set_option")

    def stat(self, name):
        """Per-category kernel time / launch count measured by the library (option "timing")."""
        out = ctypes.c_double()
        self._check(self.lib.kp_stat(self.handle, name.encode(), ctypes.byref(out)), "kp_stat")
        return out.value

    def upload_filter(self, to_filter):
        keys, off, ids = filter_csr(to_filter, self.R2)
        self._check(self.lib.kp_filter_upload(self.handle, len(keys), _np_ptr(keys), _np_ptr(off), _np_ptr(ids)),
                    "kp_filter_upload")
        self.has_filter = True

    def build_filter(self, facts):
        """facts: [F, 3] int32 rows (entity, relation, id) (Dataset.filter_facts()); the CSR is built on the device."""
        facts = np.ascontiguousarray(facts, dtype=np.int32).reshape(-1, 3)
        self._check(self.lib.kp_filter_build(self.handle, len(facts), _np_ptr(facts), self._stream()), "kp_filter_build")
        self.has_filter = True

    def download_filter(self):
        """(keys, offsets, ids) of the resident CSR as numpy arrays."""
        nk, ni = c_int64(), c_int64()
        self._check(self.lib.kp_filter_download(self.handle, ctypes.byref(nk), ctypes.byref(ni), None, None, None), "kp_filter_download")
        keys, off, ids = np.zeros(nk.value, np.int64), np.zeros(nk.value + 1, np.int64), np.zeros(ni.value, np.int32)
        self._check(self.lib.kp_filter_download(self.handle, ctypes.byref(nk), ctypes.byref(ni), _np_ptr(keys), _np_ptr(off),
                                                _np_ptr(ids)), "kp_filter_download")
        return keys, off, ids

    def all_scores(self, triples, mimic_rows=None):
        """[Q,3] int triples -> [Q, N(+1)] fp32 device tensor."""
        t = self.dev(triples, torch.int32).view(-1, 3)
        Q = t.shape[0]
        m = None if mimic_rows is None else self.dev(mimic_rows, torch.float32).view(Q, self.D)
        cols = self.N + (1 if m is not None else 0)
        out = torch.empty((Q, cols), dtype=torch.float32, device=self.device)
        self._check(self.lib.kp_all_scores(self.handle, Q, _ptr(t), _ptr(m), _ptr(out), cols, self._stream()),
                    "kp_all_scores")
        return out

    def score_triples(self, triples, mimic_rows=None):
        """[Q,3] int triples -> [Q] fp32 device tensor: the score of each triple itself (Model.score), Q rows of work."""
        t = self.dev(triples, torch.int32).view(-1, 3)
        Q = t.shape[0]
        m = None if mimic_rows is None else self.dev(mimic_rows, torch.float32).view(Q, self.D)
        out = torch.empty(Q, dtype=torch.float32, device=self.device)
        self._check(self.lib.kp_score_triples(self.handle, Q, _ptr(t), _ptr(m), _ptr(out), self._stream()), "kp_score_triples")
        return out

    def filtered_rank(self, triples, mode, mimic_rows=None, flt_off=None, flt_ids=None, counters=False):
        """Returns (target_score[Q] f32, best_score[Q] f32, rank[Q] i64[, counters[Q,4] i32]) on device."""
        t, m, fo, fi = self.dev_many([(triples, torch.int32), (mimic_rows, torch.float32), (flt_off, torch.int64), (flt_ids, torch.int32)])
        t = t.view(-1, 3)
        Q = t.shape[0]
        m = None if m is None else m.view(Q, self.D)
        if fo is not None and fi is None:
            fi = torch.zeros(1, dtype=torch.int32, device=self.device)
        ts = torch.empty(Q, dtype=torch.float32, device=self.device)
        bs = torch.empty(Q, dtype=torch.float32, device=self.device)
        rk = torch.empty(Q, dtype=torch.int64, device=self.device)
        cn = torch.empty((Q, 4), dtype=torch.int32, device=self.device) if counters else None
        self._check(self.lib.kp_filtered_rank(self.handle, Q, _ptr(t), _ptr(m), _ptr(fo), _ptr(fi), int(mode),
                                              _ptr(ts), _ptr(bs), _ptr(rk), _ptr(cn), self._stream()),
                    "kp_filtered_rank")
        return (ts, bs, rk, cn) if counters else (ts, bs, rk)

    def dp_relevance(self, preds, facts, entities, epsilon, lambd=1.0, sufficient=False):
        """Data-poisoning relevances (kp_dp_relevance): [n] fp32 device tensor."""
        p = self.dev(preds, torch.int32).view(-1, 3)
        f = self.dev(facts, torch.int32).view(-1, 3)
        e = self.dev(entities, torch.int32).view(-1)
        out = torch.empty(p.shape[0], dtype=torch.float32, device=self.device)
        self._check(self.lib.kp_dp_relevance(self.handle, p.shape[0], _ptr(p), _ptr(f), _ptr(e), float(epsilon), float(lambd),
                                             1 if sufficient else 0, _ptr(out), self._stream()), "kp_dp_relevance")
        return out

    def contract(self, queries, mode=0):
        """Diagnostic: fused score -> softmax (0) / sigmoid (1) -> contract pass; returns (m[G], l[G], O[G,D])."""
        q = self.dev(queries, torch.float32).view(-1, self.D).contiguous()
        G = q.shape[0]
        m = torch.empty(G, dtype=torch.float32, device=self.device)
        l = torch.empty(G, dtype=torch.float32, device=self.device)
        O = torch.empty((G, self.D), dtype=torch.float32, device=self.device)
        self._check(self.lib.kp_debug_contract(self.handle, G, _ptr(q), int(mode), _ptr(m), _ptr(l), _ptr(O),
                                               self._stream()), "kp_debug_contract")
        return m, l, O

    def post_train(self, hp, init_rows, row_off, rows_per_epoch, pos=None, neg=None, pos_off=None, pos_ids=None,
                   static_epochs=False, dropout_seed=0, max_rows_per_epoch=None, total_rows=None,
                   fact_off=None, facts=None, pos_idx=None, neg_code=None):
        """Run one batch of C mimic post-trainings; returns the [C, D] post-trained rows (device).
        TransE: either pos / neg or the compact tables fact_off / facts / pos_idx (uint16) / neg_code (kelpie_b200.h)."""
        staged = self.dev_many([(init_rows, torch.float32), (row_off, torch.int64), (rows_per_epoch, torch.int32),
                                (pos, torch.int32), (neg, torch.int32), (pos_off, torch.int64), (pos_ids, torch.int32),
                                (fact_off, torch.int64), (facts, torch.int32), (pos_idx, torch.uint16), (neg_code, torch.int32)])
        init = staged[0].view(-1, self.D)
        C = init.shape[0]
        out = torch.empty_like(init)
        keep = [init, out] + staged
        d = lambda i: None if staged[i] is None else _ptr(staged[i])

        if max_rows_per_epoch is None:  # host-known totals (pass them explicitly to avoid a device read)
            rpe = np.asarray(rows_per_epoch.cpu() if isinstance(rows_per_epoch, torch.Tensor) else rows_per_epoch)
            max_rows_per_epoch = int(rpe.max()) if len(rpe) else 0
        if total_rows is None:
            ro = np.asarray(row_off.cpu() if isinstance(row_off, torch.Tensor) else row_off)
            total_rows = int(ro[-1])
        b = PTBatch(C, 1 if static_epochs else 0, int(max_rows_per_epoch), 0, int(total_rows),
                    d(1), d(2), d(3), d(4), d(5), d(6), _ptr(init), _ptr(out), int(dropout_seed),
                    d(7), d(8), d(9), d(10))
        self._check(self.lib.kp_post_train_batch(self.handle, ctypes.byref(b), ctypes.byref(hp), self._stream()),
                    "kp_post_train_batch")
        self._keep = keep  # inputs must outlive the asynchronous kernels
        return out


def make_hp(kind, hp):
    """`training` dict of configs/*.json -> kp_hp, with the reference's effective values."""
    h = HP()
    h.epochs = int(hp["epochs"])
    h.batch_size = int(hp["batch_size"])
    h.beta1, h.beta2, h.eps = 0.9, 0.999, 1e-8
    h.margin = h.reg_weight = h.label_smoothing = 0.0
    h.regularizer = 0
    if kind == "TransE":  # pairwise_ranking_optimizer.py:44-47
        h.optimizer, h.lr = OPT_ADAM, float(hp["lr"])
        h.margin, h.reg_weight = float(hp["margin"]), float(hp["regularizer_weight"])
    elif kind == "ComplEx":  # multiclass_nll_optimizer.py:41-51
        name = hp.get("optimizer_name", "Adagrad")
        h.optimizer = {"Adagrad": OPT_ADAGRAD, "Adam": OPT_ADAM, "SGD": OPT_SGD}[name]
        h.lr = float(hp["lr"])
        if name == "Adam":
            h.beta1, h.beta2 = float(hp["decay1"]), float(hp["decay2"])
        if name == "Adagrad":
            h.eps = 1e-10
        h.reg_weight = float(hp["regularizer_weight"])
        h.regularizer = {"N3": 0, "N2": 1}[hp.get("regularizer_name", "N3")]  # multiclass_nll_optimizer.py:46-49
    else:  # bce_optimizer.py:165 -- Adam re-created with DEFAULT lr; the config lr is ignored
        h.optimizer, h.lr = OPT_ADAM, 1e-3
        h.label_smoothing = float(hp["label_smoothing"])
    return h


class TransEFit:
    """Full-model TransE trainer state (kp_transe_fit_*): Adam over the entity and relation tables, which are
    CUDA fp32 tensors updated in place (pairwise_ranking_optimizer.py:139-157)."""

    def __init__(self, ent, rel, norm, lr, margin, reg_weight):
        self.lib = load_library()
        if not (ent.is_cuda and rel.is_cuda and ent.is_contiguous() and rel.is_contiguous()
                and ent.dtype == torch.float32 and rel.dtype == torch.float32):
            raise RuntimeError("TransEFit needs contiguous CUDA fp32 tables (kelpie_b200 has no CPU path)")
        self.ent, self.rel = ent, rel
        self.device = ent.device
        h = c_void_p()
        rc = self.lib.kp_transe_fit_create(ent.device.index or 0, ent.shape[0], rel.shape[0], ent.shape[1], int(norm),
                                           float(lr), float(margin), float(reg_weight), _ptr(ent), _ptr(rel), ctypes.byref(h))
        if rc != 0:
            raise RuntimeError(f"kp_transe_fit_create failed ({rc}): {self.lib.kp_transe_fit_error(None).decode()}")
        self.handle = h

    def steps(self, pos, neg, step_off, want_loss=False):
        """pos / neg: [rows, 3] int32 (host or device); step_off: [n_steps + 1] row offsets (host)."""
        pos = torch.as_tensor(pos, dtype=torch.int32).to(self.device).contiguous()
        neg = torch.as_tensor(neg, dtype=torch.int32).to(self.device).contiguous()
        off = np.ascontiguousarray(step_off, dtype=np.int64)
        n = len(off) - 1
        loss = torch.zeros(max(n, 1), dtype=torch.float32, device=self.device) if want_loss else None
        rc = self.lib.kp_transe_fit_steps(self.handle, n, _np_ptr(off), _ptr(pos), _ptr(neg), _ptr(loss),
                                          c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"kp_transe_fit_steps failed ({rc}): {self.lib.kp_transe_fit_error(self.handle).decode()}")
        self._keep = (pos, neg)
        return loss

    def launches(self):
        return int(self.lib.kp_transe_fit_launches(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            torch.cuda.synchronize(self.device)
            self.lib.kp_transe_fit_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ComplExFit:
    """Full-model ComplEx trainer state (kp_complex_fit_*): 1-vs-all cross-entropy + Adagrad / Adam / SGD over the
    entity and relation tables, CUDA fp32 tensors updated in place (multiclass_nll_optimizer.py:123-135)."""

    OPTIMIZERS = {"Adagrad": 0, "Adam": 1, "SGD": 2}

    def __init__(self, ent, rel, optimizer_name, lr, decay1, decay2, reg_weight, max_batch):
        self.lib = load_library()
        if not (ent.is_cuda and rel.is_cuda and ent.is_contiguous() and rel.is_contiguous()
                and ent.dtype == torch.float32 and rel.dtype == torch.float32):
            raise RuntimeError("ComplExFit needs contiguous CUDA fp32 tables (kelpie_b200 has no CPU path)")
        self.ent, self.rel = ent, rel
        self.device = ent.device
        h = c_void_p()
        rc = self.lib.kp_complex_fit_create(ent.device.index or 0, ent.shape[0], rel.shape[0], ent.shape[1],
                                            self.OPTIMIZERS[optimizer_name], float(lr), float(decay1), float(decay2),
                                            float(reg_weight), int(max_batch), _ptr(ent), _ptr(rel), ctypes.byref(h))
        if rc != 0:
            raise RuntimeError(f"kp_complex_fit_create failed ({rc}): {self.lib.kp_complex_fit_error(None).decode()}")
        self.handle = h

    def steps(self, rows, step_off, want_loss=False):
        """rows: [total, 3] int32 (host or device), permuted; step_off: [n_steps + 1] row offsets (host)."""
        rows = torch.as_tensor(rows, dtype=torch.int32).to(self.device).contiguous()
        off = np.ascontiguousarray(step_off, dtype=np.int64)
        n = len(off) - 1
        loss = torch.zeros(max(n, 1), dtype=torch.float32, device=self.device) if want_loss else None
        rc = self.lib.kp_complex_fit_steps(self.handle, n, _np_ptr(off), _ptr(rows), _ptr(loss),
                                           c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"kp_complex_fit_steps failed ({rc}): {self.lib.kp_complex_fit_error(self.handle).decode()}")
        self._keep = rows
        return loss

    def launches(self):
        return int(self.lib.kp_complex_fit_launches(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            torch.cuda.synchronize(self.device)
            self.lib.kp_complex_fit_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ConvEParams(ctypes.Structure):
    """kp_conve_params (include/kelpie_b200.h): device pointers of every trainable tensor + batch-norm running statistics."""
    _fields_ = [(n, c_void_p) for n in ("ent", "rel", "conv_w", "conv_b", "fc_w", "fc_b",
                                         "bn1_w", "bn1_b", "bn1_mean", "bn1_var", "bn2_w", "bn2_b", "bn2_mean", "bn2_var",
                                         "bn3_w", "bn3_b", "bn3_mean", "bn3_var")] + [
        ("n_filters", c_int32), ("hidden", c_int32),
        ("drop_input", ctypes.c_float), ("drop_feature", ctypes.c_float), ("drop_hidden", ctypes.c_float)]


class ConvEFit:
    """Full-model ConvE trainer state (kp_conve_fit_*): 1-vs-all BCE with label smoothing + Adam over every parameter of
    the network, CUDA fp32 tensors updated in place (bce_optimizer.py:137-158, conve.py:133-158)."""

    def __init__(self, tensors, n_filters, hidden, dropouts, label_smoothing, max_batch, pairs, pos_off, pos_ids, seed=0):
        self.lib = load_library()
        for k, t in tensors.items():
            if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
                raise RuntimeError(f"ConvEFit needs contiguous CUDA fp32 tensors ({k}); kelpie_b200 has no CPU path")
        self.tensors = tensors
        ent, rel = tensors["ent"], tensors["rel"]
        self.device = ent.device
        prm = ConvEParams(n_filters=int(n_filters), hidden=int(hidden), drop_input=float(dropouts[0]),
                          drop_feature=float(dropouts[1]), drop_hidden=float(dropouts[2]),
                          **{k: t.data_ptr() for k, t in tensors.items()})
        self.pairs = torch.as_tensor(np.ascontiguousarray(pairs, dtype=np.int32)).to(self.device)
        self.pos_off = torch.as_tensor(np.ascontiguousarray(pos_off, dtype=np.int64)).to(self.device)
        self.pos_ids = torch.as_tensor(np.ascontiguousarray(pos_ids, dtype=np.int32)).to(self.device)
        h = c_void_p()
        rc = self.lib.kp_conve_fit_create(ent.device.index or 0, ent.shape[0], rel.shape[0], ent.shape[1], ctypes.byref(prm),
                                          float(label_smoothing), int(max_batch), len(self.pairs), _ptr(self.pairs),
                                          _ptr(self.pos_off), _ptr(self.pos_ids), int(seed), ctypes.byref(h))
        if rc != 0:
            raise RuntimeError(f"kp_conve_fit_create failed ({rc}): {self.lib.kp_conve_fit_error(None).decode()}")
        self.handle = h

    def steps(self, order, step_off, lr, want_loss=False):
        """order: [n_pairs] int32, this epoch's shuffle of the pair table; step_off: [n_steps + 1] offsets into it (host)."""
        order = torch.as_tensor(np.ascontiguousarray(order, dtype=np.int32)).to(self.device)
        off = np.ascontiguousarray(step_off, dtype=np.int64)
        n = len(off) - 1
        loss = torch.zeros(max(n, 1), dtype=torch.float32, device=self.device) if want_loss else None
        rc = self.lib.kp_conve_fit_steps(self.handle, n, _np_ptr(off), _ptr(order), float(lr), _ptr(loss),
                                         c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"kp_conve_fit_steps failed ({rc}): {self.lib.kp_conve_fit_error(self.handle).decode()}")
        self._keep = order
        return loss

    def launches(self):
        return int(self.lib.kp_conve_fit_launches(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            torch.cuda.synchronize(self.device)
            self.lib.kp_conve_fit_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
