"""Golden outputs of the verify_explanations compute flow (src/verify_explanations.py:66-262) from the
UNMODIFIED reference classes (CPU-patched): Dataset.remove/add_training_triples on a deep copy, a fresh
TransE trained by PairwiseRankingOptimizer, Model.predict_triples before / after -- the statements of
main() without its click / file-IO shell, on the 300-entity synthetic KG (dimension 64, 3 epochs).

    python tests/golden/make_golden_verify.py
"""
import copy
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refshim  # noqa: E402

refshim.install(cpu=True)

from src.data import Dataset  # noqa: E402
from src.data.dataset import MANY_TO_ONE, ONE_TO_ONE  # noqa: E402
from src.link_prediction.models import TransE  # noqa: E402
from src.link_prediction.models.transe import TransEHyperParams  # noqa: E402
from src.link_prediction.optimization import PairwiseRankingOptimizer  # noqa: E402
from src.link_prediction.optimization.pairwise_ranking_optimizer import PairwiseRankingOptimizerHyperParams  # noqa: E402

from tests.golden.make_golden import seed_all, synthetic_kg  # noqa: E402

CONFIG = dict(model_params=dict(dimension=64, norm=2),
              training=dict(batch_size=256, epochs=3, lr=0.01, margin=5, negative_triples_ratio=5, regularizer_weight=1.0))


def retrain(new_dataset, out, tag):
    new_model = TransE(dataset=new_dataset, hp=TransEHyperParams(**CONFIG["model_params"]), init_random=True)
    out[tag + "ent0"] = new_model.entity_embeddings.detach().numpy().copy()
    out[tag + "rel0"] = new_model.relation_embeddings.detach().numpy().copy()
    opt = PairwiseRankingOptimizer(model=new_model, hp=PairwiseRankingOptimizerHyperParams(**CONFIG["training"]), verbose=False)
    opt.train(training_triples=new_dataset.training_triples)
    new_model.eval()
    return new_model


if __name__ == "__main__":
    n_ent, n_rel = 300, 10
    train, valid, test = synthetic_kg(7, n_ent, n_rel, 2400, 150, 150)
    refshim.register_dataset("golden-verify", train, valid, test, n_ent, n_rel)
    dataset = Dataset("golden-verify")
    out = dict(train=train, valid=valid, test=test, n_ent=np.int64(n_ent), n_rel=np.int64(n_rel))
    seed_all(3)
    model = TransE(dataset=dataset, hp=TransEHyperParams(**CONFIG["model_params"]), init_random=True)
    model.eval()
    out["w_ent"] = model.entity_embeddings.detach().numpy().copy()
    out["w_rel"] = model.relation_embeddings.detach().numpy().copy()
    meta = dict(config=CONFIG)

    preds = []
    for s, p, o in dataset.testing_triples:
        if 4 <= len(dataset.entity_to_training_triples[s]) <= 14 and (int(s), int(p), int(o)) not in preds:
            preds.append((int(s), int(p), int(o)))
        if len(preds) == 4:
            break
    rules = {p: [tuple(int(x) for x in t) for t in dataset.entity_to_training_triples[p[0]][:2]] for p in preds}

    # ---- necessary (verify_explanations.py:197-262)
    seed_all(31)
    nec = preds[:3]
    new_dataset = copy.deepcopy(dataset)
    new_dataset.remove_training_triples([t for p in nec for t in rules[p]])
    results = model.predict_triples(np.array(nec))
    new_model = retrain(new_dataset, out, "nec_")
    new_results = new_model.predict_triples(np.array(nec))
    out["nec_scores"] = np.array([[r["score"]["tail"], n["score"]["tail"]] for r, n in zip(results, new_results)], dtype=np.float64)
    out["nec_ranks"] = np.array([[r["rank"]["tail"], n["rank"]["tail"]] for r, n in zip(results, new_results)], dtype=np.int64)
    out["nec_ent"] = new_model.entity_embeddings.detach().numpy().copy()
    meta["necessary"] = dict(seed=31, preds=nec, rules={str(list(p)): rules[p] for p in nec})

    # ---- sufficient (verify_explanations.py:66-195)
    seed_all(32)
    suf = preds[2:4]
    rng = np.random.default_rng(5)
    entities = {p: [int(e) for e in rng.choice([e for e in range(n_ent) if e != p[0]], size=3, replace=False)] for p in suf}
    to_add, to_convert = [], []
    for pred in suf:
        for e in entities[pred]:
            to_convert.append(Dataset.replace_entity_in_triple(pred, pred[0], e))
            to_add.extend(Dataset.replace_entity_in_triples(rules[pred], pred[0], e))
    new_dataset = copy.deepcopy(dataset)
    for s, p, o in to_add:
        if new_dataset.relation_to_type[p] in [MANY_TO_ONE, ONE_TO_ONE]:
            for existing_o in new_dataset.train_to_filter[(s, p)]:
                new_dataset.remove_training_triple((s, p, existing_o))
    new_dataset.add_training_triples(to_add)
    results = model.predict_triples(np.array(to_convert))
    new_model = retrain(new_dataset, out, "suf_")
    new_results = new_model.predict_triples(np.array(to_convert))
    out["suf_scores"] = np.array([[r["score"]["tail"], n["score"]["tail"]] for r, n in zip(results, new_results)], dtype=np.float64)
    out["suf_ranks"] = np.array([[r["rank"]["tail"], n["rank"]["tail"]] for r, n in zip(results, new_results)], dtype=np.int64)
    out["suf_n_train"] = np.int64(len(new_dataset.training_triples))
    meta["sufficient"] = dict(seed=32, preds=suf, rules={str(list(p)): rules[p] for p in suf},
                              entities={str(list(p)): entities[p] for p in suf}, to_convert=[list(map(int, t)) for t in to_convert])
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "verify_small.npz")
    np.savez_compressed(path, **out)
    print("->", path, os.path.getsize(path), "bytes", out["nec_ranks"].tolist(), out["suf_ranks"].tolist())
