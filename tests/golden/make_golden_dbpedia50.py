"""Golden vectors at BASELINE.json's own configs on the REAL DBpedia50 dataset, produced by the
UNMODIFIED reference (CPU-patched) in the build container:

    python tests/golden/make_golden_dbpedia50.py

* configs/TransE_DBpedia50_explanation.json  (configs[0]: dim 256, L2, 65 epochs, Adam 0.01, margin 5)
  necessary mode: 3 test predictions x up to 6 single-fact candidates + one 2-fact candidate
* configs/ComplEx_DBpedia50_explanation.json (configs[1]: dim 200, Adagrad 0.043, 43 epochs)
  necessary mode: 2 predictions x 4 candidates; sufficient mode: 1 prediction, 3 conversion entities
  chosen by the reference's select_entities_to_convert over all 24 620 entities, 2 candidates

Trained checkpoints are offline (figshare), so the weights are drawn from a seeded CPU generator by the
recipe in `weights()` below and assigned into the reference's model; the tests regenerate the same tensors
(24 620 x 400 floats do not fit a fixture) and check their checksum.  Stored: predictions, candidate
rules, seeds, the fact order the reference's sets produced, and the reference's outputs (relevances,
(score, rank, best score) of every post-training, post-trained mimic rows, conversion entities).
"""
import json
import os
import random
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refshim  # noqa: E402

refshim.install(cpu=True)

from src.data import Dataset  # noqa: E402
from src.link_prediction.models import ComplEx, TransE  # noqa: E402
from src.link_prediction.models.complex import ComplExHyperParams  # noqa: E402
from src.link_prediction.models.transe import TransEHyperParams  # noqa: E402
from src.relevance_engines import NecessaryPostTrainingEngine, SufficientPostTrainingEngine  # noqa: E402

from tests.golden.make_golden import pack_trace, seed_all, traced  # noqa: E402


def weights(kind, n_ent, n_rel2, row):
    """Seeded stand-in for a trained checkpoint (same recipe in tests/test_gpu_dbpedia50.py)."""
    g = torch.Generator().manual_seed(20240 + (0 if kind == "TransE" else 1))
    scale = 0.35 if kind == "TransE" else 0.25
    ent = torch.randn(n_ent, row, generator=g) * scale
    rel = torch.randn(n_rel2, row, generator=g) * scale
    return ent, rel


CONFIGS = {
    "TransE": dict(cls=TransE, hpc=TransEHyperParams, params=dict(dimension=256, norm=2),
                   hp=dict(batch_size=2048, epochs=65, lr=0.01, margin=5, negative_triples_ratio=5, regularizer_weight=1.0)),
    "ComplEx": dict(cls=ComplEx, hpc=ComplExHyperParams, params=dict(dimension=200, init_scale=1e-3),
                    hp=dict(optimizer_name="Adagrad", batch_size=512, epochs=43, lr=0.043, decay1=0.9, decay2=0.999,
                            regularizer_name="N3", regularizer_weight=0)),
}


def generate(kind, ds):
    cfg = CONFIGS[kind]
    seed_all(42)
    model = cfg["cls"](ds, cfg["hpc"](**cfg["params"]), init_random=True)
    ent, rel = weights(kind, ds.num_entities, 2 * ds.num_relations, model.entity_embeddings.shape[1])
    with torch.no_grad():
        model.entity_embeddings.copy_(ent)
        model.relation_embeddings.copy_(rel)
    model.eval()
    out = {"w_checksum": np.array([float(ent.double().sum()), float(rel.double().sum()), float(ent.double().abs().sum())])}
    meta = dict(kind=kind, params=cfg["params"], hp=cfg["hp"], cases=[])
    fact_order = {}

    preds = []
    for s, p, o in ds.testing_triples:
        deg = len(ds.entity_to_training_triples[s])
        if 4 <= deg <= 9 and (int(s), int(p), int(o)) not in preds:
            preds.append((int(s), int(p), int(o)))
        if len(preds) == 4:
            break

    n_nec = 3 if kind == "TransE" else 2
    eng = NecessaryPostTrainingEngine(model, ds, cfg["hp"])
    for pi, pred in enumerate(preds[:n_nec]):
        facts = [tuple(int(x) for x in t) for t in ds.entity_to_training_triples[pred[0]]]
        fact_order[pred[0]] = facts
        if kind == "TransE":
            rules = [[f] for f in facts[:6]] + [[facts[0], facts[1]]]
        else:
            rules = [[facts[0]], [facts[1]], [facts[2]], [facts[0], facts[3]]]
        seed_all(100 + pi)
        eng.set_cache()
        trace = traced(eng)
        t0 = time.time()
        rels = [eng.compute_relevance(pred, r) for r in rules]
        print(kind, "necessary", pred, len(rules), "candidates", f"{time.time() - t0:.1f} s", np.round(rels, 4))
        tag = f"nec{pi}_"
        pack_trace(trace, tag, out)
        out[tag + "relevance"] = np.array(rels, dtype=np.float64)
        meta["cases"].append(dict(tag=tag, mode="necessary", pred=pred, rules=rules, seed=100 + pi))

    if kind == "ComplEx":
        pred = preds[3]
        facts = [tuple(int(x) for x in t) for t in ds.entity_to_training_triples[pred[0]]]
        fact_order[pred[0]] = facts
        eng = SufficientPostTrainingEngine(model, ds, cfg["hp"])
        seed_all(200)
        eng.set_cache()
        t0 = time.time()
        eng.select_entities_to_convert(pred, 3, 200)
        conv = [int(e) for e in eng.entities_to_convert]
        print("select_entities_to_convert", conv, f"{time.time() - t0:.1f} s")
        for e in conv:
            fact_order[e] = [tuple(int(x) for x in t) for t in ds.entity_to_training_triples[e]]
        rules = [[facts[0]], [facts[1], facts[2]]]
        trace = traced(eng)
        t0 = time.time()
        rels = [eng.compute_relevance(pred, r) for r in rules]
        print(kind, "sufficient", pred, f"{time.time() - t0:.1f} s", np.round(rels, 4))
        pack_trace(trace, "suf0_", out)
        out["suf0_relevance"] = np.array(rels, dtype=np.float64)
        meta["cases"].append(dict(tag="suf0_", mode="sufficient", pred=pred, rules=rules, seed=200, entities_to_convert=conv))

    meta["fact_order"] = {str(k): v for k, v in fact_order.items()}
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, f"dbpedia50_{kind.lower()}.npz")
    np.savez_compressed(path, **out)
    print(kind, "->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    cwd = os.getcwd()
    os.chdir(refshim.REFERENCE_ROOT)
    try:
        ds = Dataset("DBpedia50")
    finally:
        os.chdir(cwd)
    for k in sys.argv[1:] or ("TransE", "ComplEx"):
        generate(k, ds)
