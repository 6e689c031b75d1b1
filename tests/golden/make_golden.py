"""Generate tests/golden/*.npz by running the UNMODIFIED reference (CPU-patched) here.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For each of TransE / ComplEx / ConvE it builds a small synthetic KG, random weights
(the reference's own init, seed 42), runs the reference's Necessary- and
SufficientPostTrainingEngine, Model.all_scores, Model.predict_triples and
RelevanceEngine.select_entities_to_convert, and stores inputs (triples, weights,
hyper-parameters, seeds, fact order) and the reference's outputs (relevances, ranks,
target scores, post-trained mimic rows).  tests/test_oracle_golden.py replays the
same seeds through oracle/kelpie_oracle.py and the CUDA path and compares.

It also stores the id-mapped DBpedia50 triples (labels dropped; ids = sorted labels of
the training split, unseen valid/test rows dropped) so the GPU box can run the
DBpedia50-shaped configs without /root/reference.
"""
import json
import os
import random
import sys

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
from src.relevance_engines import (  # noqa: E402
    NecessaryPostTrainingEngine,
    SufficientPostTrainingEngine,
)


def seed_all(seed):
    np.random.seed(seed)
    torch.manual_seed(seed)
    random.seed(seed)


def synthetic_kg(seed, n_ent, n_rel, n_train, n_valid, n_test):
    rng = np.random.default_rng(seed)

    def draw(n):
        # mildly skewed degrees, a few popular relations
        s = (rng.random(n) ** 1.6 * n_ent).astype(np.int64)
        o = (rng.random(n) ** 1.6 * n_ent).astype(np.int64)
        p = (rng.random(n) ** 1.5 * n_rel).astype(np.int64)
        return np.stack([s, p, o], 1)

    train = draw(n_train)
    train[:n_ent, 0] = rng.permutation(n_ent)  # every entity occurs in training
    train[:n_rel, 1] = np.arange(n_rel)
    train = np.unique(train, axis=0)
    rng.shuffle(train)
    return train, draw(n_valid), draw(n_test)


MODELS = {
    "TransE": dict(
        params=dict(dimension=64, norm=2),
        hp=dict(batch_size=2048, epochs=12, lr=0.01, margin=5, negative_triples_ratio=5,
                regularizer_weight=1.0),
    ),
    "ComplEx": dict(
        params=dict(dimension=32, init_scale=1e-3),
        hp=dict(optimizer_name="Adagrad", batch_size=512, epochs=10, lr=0.043, decay1=0.9,
                decay2=0.999, regularizer_name="N3", regularizer_weight=0),
    ),
    "ConvE": dict(
        params=dict(dimension=80, input_dropout_rate=0.0, feature_map_dropout_rate=0.0,
                    hidden_dropout_rate=0.0, hidden_layer_size=2432),
        hp=dict(batch_size=512, label_smoothing=0.1, lr=0.018, decay=0.995, epochs=8),
    ),
}


def build_model(kind, ds):
    p = MODELS[kind]["params"]
    if kind == "TransE":
        m = TransE(ds, TransEHyperParams(**p))
    elif kind == "ComplEx":
        m = ComplEx(ds, ComplExHyperParams(**p))
        with torch.no_grad():  # init_scale 1e-3 leaves logits ~0: widen so ranks are informative
            m.entity_embeddings.copy_(torch.randn_like(m.entity_embeddings) * 0.3)
            m.relation_embeddings.copy_(torch.randn_like(m.relation_embeddings) * 0.3)
    else:
        m = ConvE(ds, ConvEHyperParams(**p))
        with torch.no_grad():  # non-trivial eval-mode batch-norm statistics
            for bn in (m.batch_norm_1, m.batch_norm_2, m.batch_norm_3):
                bn.running_mean.copy_(torch.randn_like(bn.running_mean) * 0.1)
                bn.running_var.copy_(torch.rand_like(bn.running_var) * 0.5 + 0.75)
                bn.weight.copy_(torch.rand_like(bn.weight) * 0.5 + 0.75)
                bn.bias.copy_(torch.randn_like(bn.bias) * 0.1)
    m.eval()
    return m


def weights_of(kind, m):
    out = {
        "ent": m.entity_embeddings.detach().numpy().copy(),
        "rel": m.relation_embeddings.detach().numpy().copy(),
    }
    if kind == "ConvE":
        out.update(
            conv_w=m.convolutional_layer.weight.detach().numpy().copy(),
            conv_b=m.convolutional_layer.bias.detach().numpy().copy(),
            fc_w=m.hidden_layer.weight.detach().numpy().copy(),
            fc_b=m.hidden_layer.bias.detach().numpy().copy(),
        )
        for i, bn in enumerate((m.batch_norm_1, m.batch_norm_2, m.batch_norm_3), 1):
            out[f"bn{i}_w"] = bn.weight.detach().numpy().copy()
            out[f"bn{i}_b"] = bn.bias.detach().numpy().copy()
            out[f"bn{i}_mean"] = bn.running_mean.numpy().copy()
            out[f"bn{i}_var"] = bn.running_var.numpy().copy()
    return out


def traced(engine):
    """Record (mimic row before, after) of every post-training and every rank result."""
    trace = []
    orig_pt, orig_res = engine.post_train, engine.get_triple_results

    def post_train(model, triples):
        before = model.kelpie_entity_emb.detach().clone().numpy()
        out = orig_pt(model=model, triples=triples)
        after = model.kelpie_entity_emb.detach().clone().numpy()
        trace.append(["pt", before, after, None])
        return out

    def get_triple_results(model, triple):
        res = orig_res(model, triple)
        trace[-1][3] = (float(res["target_score"]), int(res["target_rank"]), float(res["best_score"]))
        return res

    engine.post_train = post_train
    engine.get_triple_results = get_triple_results
    return trace


def pack_trace(trace, prefix, out):
    out[prefix + "n"] = np.int64(len(trace))
    for i, (_, before, after, res) in enumerate(trace):
        out[f"{prefix}{i}_init"] = before
        out[f"{prefix}{i}_final"] = after
        out[f"{prefix}{i}_res"] = np.array(res, dtype=np.float64)


def generate(kind, seed=42):
    n_ent, n_rel = 300, 10
    train, valid, test = synthetic_kg(7, n_ent, n_rel, 2400, 150, 150)
    name = f"golden-{kind}"
    refshim.register_dataset(name, train, valid, test, n_ent, n_rel)
    ds = Dataset(name)
    seed_all(seed)
    model = build_model(kind, ds)
    hp = MODELS[kind]["hp"]
    out = dict(train=train, valid=valid, test=test, n_ent=np.int64(n_ent), n_rel=np.int64(n_rel))
    out.update({"w_" + k: v for k, v in weights_of(kind, model).items()})
    meta = dict(kind=kind, params=MODELS[kind]["params"], hp=hp, seed=seed, cases=[])

    # --- predictions to explain: test triples whose head has 4..14 training facts
    preds = []
    for s, p, o in ds.testing_triples:
        deg = len(ds.entity_to_training_triples[s])
        if 4 <= deg <= 14 and (s, p, o) not in preds:
            preds.append((int(s), int(p), int(o)))
        if len(preds) == 3:
            break
    fact_order = {}

    # --- necessary mode
    eng = NecessaryPostTrainingEngine(model, ds, hp)
    for pi, pred in enumerate(preds[:2]):
        facts = [tuple(int(x) for x in t) for t in ds.entity_to_training_triples[pred[0]]]
        fact_order[pred[0]] = facts
        rules = [[facts[0]], [facts[1]], [facts[2]], [facts[0], facts[3]]]
        seed_all(seed + 1 + pi)
        eng.set_cache()
        trace = traced(eng)
        rels = [eng.compute_relevance(pred, r) for r in rules]
        tag = f"nec{pi}_"
        pack_trace(trace, tag, out)
        out[tag + "relevance"] = np.array(rels, dtype=np.float64)
        meta["cases"].append(dict(tag=tag, mode="necessary", pred=pred, rules=rules, seed=seed + 1 + pi))

    # --- sufficient mode
    pred = preds[2]
    facts = [tuple(int(x) for x in t) for t in ds.entity_to_training_triples[pred[0]]]
    fact_order[pred[0]] = facts
    eng = SufficientPostTrainingEngine(model, ds, hp)
    seed_all(seed + 10)
    eng.set_cache()
    eng.select_entities_to_convert(pred, 3, 200)
    conv = [int(e) for e in eng.entities_to_convert]
    for e in conv:
        fact_order[e] = [tuple(int(x) for x in t) for t in ds.entity_to_training_triples[e]]
    rules = [[facts[0]], [facts[1], facts[2]]]
    trace = traced(eng)
    rels = [eng.compute_relevance(pred, r) for r in rules]
    pack_trace(trace, "suf0_", out)
    out["suf0_relevance"] = np.array(rels, dtype=np.float64)
    meta["cases"].append(
        dict(tag="suf0_", mode="sufficient", pred=pred, rules=rules, seed=seed + 10, entities_to_convert=conv)
    )

    # --- convertible entities (before random.sample): k larger than the pool returns all
    seed_all(seed + 20)
    eng.select_entities_to_convert(pred, 10 ** 6, 200)
    out["convertible"] = np.array(sorted(int(e) for e in eng.entities_to_convert), dtype=np.int64)
    meta["convertible_pred"] = pred

    # --- all_scores / predict_triples
    q = ds.testing_triples[:6].copy()
    with torch.no_grad():
        out["all_scores_q"] = q
        out["all_scores"] = model.all_scores(q).detach().numpy().copy()
    pt = ds.testing_triples[:24].copy()
    res = model.predict_triples(pt)
    out["predict_q"] = pt
    out["predict_scores"] = np.array([[r["score"]["tail"], r["score"]["head"]] for r in res], dtype=np.float64)
    out["predict_ranks"] = np.array([[r["rank"]["tail"], r["rank"]["head"]] for r in res], dtype=np.int64)

    meta["fact_order"] = {str(k): v for k, v in fact_order.items()}
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, f"{kind.lower()}_small.npz")
    np.savez_compressed(path, **out)
    print(kind, "->", path, os.path.getsize(path), "bytes")


def dbpedia50_ids():
    cwd = os.getcwd()
    os.chdir(refshim.REFERENCE_ROOT)  # the reference resolves data/ relative to its root
    try:
        from src import DBPEDIA50_PATH

        d = refshim._get_dataset(
            training=DBPEDIA50_PATH / "train.txt",
            testing=DBPEDIA50_PATH / "test.txt",
            validation=DBPEDIA50_PATH / "valid.txt",
        )
    finally:
        os.chdir(cwd)
    path = os.path.join(HERE, "dbpedia50_ids.npz")
    np.savez_compressed(
        path,
        train=d.training.mapped_triples.numpy().astype(np.int32),
        valid=d.validation.mapped_triples.numpy().astype(np.int32),
        test=d.testing.mapped_triples.numpy().astype(np.int32),
        n_ent=np.int64(d.num_entities),
        n_rel=np.int64(d.num_relations),
    )
    print("DBpedia50 ->", path, os.path.getsize(path), "bytes", d.num_entities, d.num_relations)


if __name__ == "__main__":
    for k in ("TransE", "ComplEx", "ConvE"):
        generate(k)
    dbpedia50_ids()
