"""TEST INFRASTRUCTURE ONLY -- import shim for running the UNMODIFIED reference on CPU.

Only `tests/golden/make_golden.py` uses this file, and only in the build container where
`/root/reference` is mounted.  Nothing in the product (`kelpie_b200/`) may import it.

What it does:
  * registers stand-ins for the three third-party imports the reference needs but that
    are absent from this image: `pykeen.datasets.get_dataset` (dataset.py:9),
    `optuna` (pairwise_ranking_optimizer.py:1 ...), `bispy` (bisimulation.py:3);
    PyKEEN supplies no arithmetic -- only label->id mapping and the three id-triple
    tensors (dataset.py:21-25,97) -- so the stand-in is a TSV reader / in-memory holder;
  * redirects the reference's hard-coded `.cuda()` calls to the CPU.  `Tensor.cuda()`
    must return a COPY: KelpieComplEx scales its parameter in place (complex.py:155-157)
    and would otherwise corrupt the shared `init_tensor`;
  * puts `/root/reference` on `sys.path` so `import src...` resolves to the reference.
"""
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("KELPIE_REFERENCE_ROOT", "/root/reference")

_REGISTERED = {}


class _TriplesFactory:
    def __init__(self, mapped):
        self.mapped_triples = torch.as_tensor(np.asarray(mapped, dtype=np.int64))


class _PykeenLikeDataset:
    """The attributes dataset.py reads from a PyKEEN dataset (dataset.py:97-190)."""

    def __init__(self, train, valid, test, entity_to_id, relation_to_id):
        self.training = _TriplesFactory(train)
        self.validation = _TriplesFactory(valid)
        self.testing = _TriplesFactory(test)
        self.entity_to_id = entity_to_id
        self.relation_to_id = relation_to_id
        self.num_entities = len(entity_to_id)
        self.num_relations = len(relation_to_id)


def register_dataset(name, train, valid, test, num_entities, num_relations):
    """Register an in-memory id-triple dataset under `name` for `Dataset(name)`."""
    e2i = {f"e{i}": i for i in range(num_entities)}
    r2i = {f"r{i}": i for i in range(num_relations)}
    _REGISTERED[name] = _PykeenLikeDataset(train, valid, test, e2i, r2i)


def _read_tsv(path):
    rows = []
    with open(path, encoding="utf-8") as f:
        for line in f:
            parts = line.rstrip("\n").split("\t")
            if len(parts) == 3:
                rows.append(parts)
    return rows


def _get_dataset(dataset=None, training=None, testing=None, validation=None, **_):
    if dataset is not None:
        return _REGISTERED[dataset]
    tr, va, te = _read_tsv(training), _read_tsv(validation), _read_tsv(testing)
    ents = sorted({h for h, _, _ in tr} | {t for _, _, t in tr})
    rels = sorted({r for _, r, _ in tr})
    e2i = {e: i for i, e in enumerate(ents)}
    r2i = {r: i for i, r in enumerate(rels)}

    def to_ids(rows):
        out = [
            (e2i[h], r2i[r], e2i[t])
            for h, r, t in rows
            if h in e2i and t in e2i and r in r2i
        ]
        return np.array(out, dtype=np.int64).reshape(-1, 3)

    return _PykeenLikeDataset(to_ids(tr), to_ids(va), to_ids(te), e2i, r2i)


def _install_modules():
    pykeen = types.ModuleType("pykeen")
    datasets = types.ModuleType("pykeen.datasets")
    datasets.get_dataset = _get_dataset
    pykeen.datasets = datasets
    sys.modules.setdefault("pykeen", pykeen)
    sys.modules.setdefault("pykeen.datasets", datasets)

    optuna = types.ModuleType("optuna")
    exceptions = types.ModuleType("optuna.exceptions")

    class TrialPruned(Exception):
        pass

    exceptions.TrialPruned = TrialPruned
    optuna.exceptions = exceptions
    sys.modules.setdefault("optuna", optuna)
    sys.modules.setdefault("optuna.exceptions", exceptions)

    bispy = types.ModuleType("bispy")

    def compute_maximum_bisimulation(*a, **k):
        raise NotImplementedError("bispy is absent; summarisation is out of scope")

    bispy.compute_maximum_bisimulation = compute_maximum_bisimulation
    sys.modules.setdefault("bispy", bispy)


def _patch_torch_for_cpu():
    if getattr(torch, "_kelpie_cpu_patched", False):
        return
    torch.Tensor.cuda = lambda self, *a, **k: self.clone()
    torch.nn.Module.cuda = lambda self, *a, **k: self
    _orig_to = torch.nn.Module.to

    def _to(self, *a, **k):
        a = tuple(x for x in a if x != "cuda")
        k.pop("device", None)
        return _orig_to(self, *a, **k) if (a or k) else self

    torch.nn.Module.to = _to

    def _strip_device(fn):
        def wrapped(*a, **k):
            if k.get("device") == "cuda":
                k.pop("device")
            return fn(*a, **k)

        return wrapped

    torch.zeros = _strip_device(torch.zeros)
    torch.tensor = _strip_device(torch.tensor)
    torch.cuda.get_rng_state = lambda *a, **k: None
    torch.cuda.set_rng_state = lambda *a, **k: None
    torch._kelpie_cpu_patched = True


def install(cpu=True):
    """Make `import src` resolve to the unmodified reference; optionally run it on CPU."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_modules()
    if cpu:
        _patch_torch_for_cpu()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
