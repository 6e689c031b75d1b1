"""Golden vectors at BASELINE.json's own configs on the REAL DBpedia50 dataset, produced by the
UNMODIFIED reference (CPU-patched) in the build container:

    python tests/golden/make_golden_dbpedia50.py

* configs/TransE_DBpedia50_explanation.json  (configs[0]: dim 256, L2, 65 epochs, Adam 0.01, margin 5)
  necessary mode: 3 test predictions x up to 6 single-fact candidates + one 2-fact candidate
* configs/ComplEx_DBpedia50_explanation.json (configs[1]: dim 200, Adagrad 0.043, 43 epochs)
  necessary mode: 2 predictions x 4 candidates; sufficient mode: 1 prediction, 3 conversion entities
  chosen by the reference's select_entities_to_convert over all 24 620 entities, 2 candidates
* configs/ConvE_DBpedia50_explanation.json (configs[2]'s arithmetic shape on real data: dim 200 = 20 x 10, hidden layer
  9728, label smoothing 0.1, 69 epochs, the optimiser's default Adam lr 1e-3, all dropout rates 0)
  necessary mode: 2 predictions x 3 candidates

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
from src.link_prediction.models import ComplEx, ConvE, TransE  # noqa: E402
from src.link_prediction.models.complex import ComplExHyperParams  # noqa: E402
from src.link_prediction.models.conve import ConvEHyperParams  # noqa: E402
from src.link_prediction.models.transe import TransEHyperParams  # noqa: E402
from src.relevance_engines import NecessaryPostTrainingEngine, SufficientPostTrainingEngine  # noqa: E402

from tests.golden.make_golden import pack_trace, seed_all, traced  # noqa: E402


def weights(kind, n_ent, n_rel2, row):
    """Seeded stand-in for a trained checkpoint (same recipe in tests/test_gpu_dbpedia50.py)."""
    g = torch.Generator().manual_seed(20240 + {"TransE": 0, "ComplEx": 1, "ConvE": 2}[kind])
    scale = 0.35 if kind == "TransE" else 0.25
    ent = torch.randn(n_ent, row, generator=g) * scale
    rel = torch.randn(n_rel2, row, generator=g) * scale
    return ent, rel


def conve_network(dim, hidden):
    """Seeded frozen ConvE network (conv, Linear, three eval-mode batch norms); same recipe in the test."""
    g = torch.Generator().manual_seed(20250)
    net = dict(conv_w=torch.randn(32, 1, 3, 3, generator=g) * 0.3, conv_b=torch.randn(32, generator=g) * 0.1,
               fc_w=torch.randn(dim, hidden, generator=g) * (1.0 / hidden) ** 0.5, fc_b=torch.randn(dim, generator=g) * 0.1)
    for i, n in ((1, 1), (2, 32), (3, dim)):
        net[f"bn{i}_w"] = torch.rand(n, generator=g) * 0.5 + 0.75
        net[f"bn{i}_b"] = torch.randn(n, generator=g) * 0.1
        net[f"bn{i}_mean"] = torch.randn(n, generator=g) * 0.1
        net[f"bn{i}_var"] = torch.rand(n, generator=g) * 0.5 + 0.75
    return net


def load_conve_network(m, net):
    with torch.no_grad():
        m.convolutional_layer.weight.copy_(net["conv_w"])
        m.convolutional_layer.bias.copy_(net["conv_b"])
        m.hidden_layer.weight.copy_(net["fc_w"])
        m.hidden_layer.bias.copy_(net["fc_b"])
        for i, bn in enumerate((m.batch_norm_1, m.batch_norm_2, m.batch_norm_3), 1):
            bn.weight.copy_(net[f"bn{i}_w"])
            bn.bias.copy_(net[f"bn{i}_b"])
            bn.running_mean.copy_(net[f"bn{i}_mean"])
            bn.running_var.copy_(net[f"bn{i}_var"])


CONFIGS = {
    "TransE": dict(cls=TransE, hpc=TransEHyperParams, params=dict(dimension=256, norm=2),
                   hp=dict(batch_size=2048, epochs=65, lr=0.01, margin=5, negative_triples_ratio=5, regularizer_weight=1.0)),
    "ComplEx": dict(cls=ComplEx, hpc=ComplExHyperParams, params=dict(dimension=200, init_scale=1e-3),
                    hp=dict(optimizer_name="Adagrad", batch_size=512, epochs=43, lr=0.043, decay1=0.9, decay2=0.999,
                            regularizer_name="N3", regularizer_weight=0)),
    "ConvE": dict(cls=ConvE, hpc=ConvEHyperParams,
                  params=dict(dimension=200, input_dropout_rate=0.0, feature_map_dropout_rate=0.0, hidden_dropout_rate=0.0,
                              hidden_layer_size=9728),
                  hp=dict(batch_size=512, label_smoothing=0.1, lr=0.018, decay=0.995, epochs=69)),
}


def generate(kind, ds):
    cfg = CONFIGS[kind]
    seed_all(42)
    model = cfg["cls"](ds, cfg["hpc"](**cfg["params"]), init_random=True) if kind != "ConvE" else cfg["cls"](ds, cfg["hpc"](**cfg["params"]))
    ent, rel = weights(kind, ds.num_entities, 2 * ds.num_relations, model.entity_embeddings.shape[1])
    with torch.no_grad():
        model.entity_embeddings.copy_(ent)
        model.relation_embeddings.copy_(rel)
    out = {"w_checksum": np.array([float(ent.double().sum()), float(rel.double().sum()), float(ent.double().abs().sum())])}
    if kind == "ConvE":
        net = conve_network(cfg["params"]["dimension"], cfg["params"]["hidden_layer_size"])
        load_conve_network(model, net)
        out["net_checksum"] = np.array([float(v.double().sum()) for v in net.values()])
    model.eval()
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
    if kind == "ConvE":
        seed_all(7)  # decouple the predictions' RNG history from the model constructor's
    eng = NecessaryPostTrainingEngine(model, ds, cfg["hp"])
    for pi, pred in enumerate(preds[:n_nec]):
        facts = [tuple(int(x) for x in t) for t in ds.entity_to_training_triples[pred[0]]]
        fact_order[pred[0]] = facts
        if kind == "TransE":
            rules = [[f] for f in facts[:6]] + [[facts[0], facts[1]]]
        elif kind == "ConvE":
            rules = [[facts[0]], [facts[1]], [facts[0], facts[2]]]
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
