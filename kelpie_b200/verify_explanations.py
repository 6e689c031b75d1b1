"""End-to-end verification of explanations (src/verify_explanations.py:66-270) as a function: edit the
training set with the best rule of every prediction, retrain the model from scratch, and compare the
filtered tail rank / score of the predictions (necessary) or of their conversions (sufficient) before
and after.  The CLI, file IO and label printing of the reference stay with it; this module is the compute
flow behind them -- `Model.predict_triples` on the rank kernels, `Optimizer.train` on the device trainer
(TransE) -- and returns the reference's `evaluations` list with ids instead of labels.
"""
import copy

import numpy as np

from .data import Dataset
from .data.names import MANY_TO_ONE, ONE_TO_ONE
from .link_prediction import MODEL_REGISTRY


def _predict(model, triples, batch_size):
    """verify_explanations.py:120-133,146-160: TransE predicts in batches, the others in one call."""
    triples = [tuple(int(x) for x in t) for t in triples]
    if len(triples) > batch_size and model.name == "TransE":
        res = []
        for b0 in range(0, len(triples), batch_size):
            res += model.predict_triples(np.array(triples[b0:b0 + batch_size]))
    else:
        res = model.predict_triples(np.array(triples))
    return dict(zip(triples, res))


def _retrain(model, new_dataset, model_config, model_factory):
    """verify_explanations.py:138-144 / :226-232: a fresh model on the edited dataset, trained with the config."""
    name = model.name
    model_class, optimizer_class = MODEL_REGISTRY[name]["class"], MODEL_REGISTRY[name]["optimizer"]
    if model_factory is not None:
        new_model = model_factory(new_dataset)
    else:
        new_model = model_class(new_dataset, model_class.get_hyperparams_class()(**model_config["model_params"]), init_random=True)
    hp = optimizer_class.get_hyperparams_class()(**model_config["training"])
    optimizer_class(model=new_model, hp=hp, verbose=False).train(training_triples=new_dataset.training_triples)
    new_model.eval()
    return new_model


def verify_necessary(model, dataset, pred_to_rule, model_config, model_factory=None):
    """pred_to_rule: {(s, p, o): [facts of the best rule]} -> [{triple_to_explain, rule, score, rank, new_score, new_rank}]
    (verify_explanations.py:197-262)."""
    preds = [tuple(int(x) for x in p) for p in pred_to_rule]
    to_remove = [tuple(int(x) for x in t) for p in pred_to_rule for t in pred_to_rule[p]]
    new_dataset = copy.deepcopy(dataset)
    new_dataset.remove_training_triples(to_remove)
    results = _predict(model, preds, len(preds) + 1)
    new_model = _retrain(model, new_dataset, model_config, model_factory)
    new_results = _predict(new_model, preds, len(preds) + 1)
    out = []
    for p in preds:
        r, n = results[p], new_results[p]
        out.append({"triple_to_explain": p, "rule": [tuple(int(x) for x in t) for t in pred_to_rule[p]],
                    "score": str(r["score"]["tail"]), "rank": str(r["rank"]["tail"]),
                    "new_score": str(n["score"]["tail"]), "new_rank": str(n["rank"]["tail"])})
    return out, new_model


def verify_sufficient(model, dataset, pred_to_rule, pred_to_entities, model_config, model_factory=None):
    """verify_explanations.py:66-195: the best rule of every prediction is added to each of its conversion
    entities (existing objects of *-to-one relations are removed first), the model is retrained, and the
    converted predictions <e, p, o> are ranked before and after."""
    preds = [tuple(int(x) for x in p) for p in pred_to_rule]
    to_add, to_convert, convert_set, convert_to_added = [], [], {}, {}
    for pred in preds:
        s = pred[0]
        cur = []
        for e in pred_to_entities[pred]:
            c = Dataset.replace_entity_in_triple(pred, s, int(e))
            cur.append(c)
            added = Dataset.replace_entity_in_triples([tuple(int(x) for x in t) for t in pred_to_rule[pred]], s, int(e))
            to_add.extend(added)
            convert_to_added[c] = added
        to_convert.extend(cur)
        convert_set[pred] = cur
    new_dataset = copy.deepcopy(dataset)
    for s, p, o in to_add:
        if new_dataset.relation_to_type[p] in (MANY_TO_ONE, ONE_TO_ONE):
            for existing_o in list(new_dataset.train_to_filter[(s, p)]):
                new_dataset.remove_training_triple((s, p, existing_o))
    new_dataset.add_training_triples(to_add)
    results = _predict(model, to_convert, 256)
    new_model = _retrain(model, new_dataset, model_config, model_factory)
    new_results = _predict(new_model, to_convert, 64)
    out = []
    for pred in preds:
        conv = []
        for c in convert_set[pred]:
            r, n = results[c], new_results[c]
            conv.append({"triples_to_add": convert_to_added[c], "score": str(r["score"]["tail"]), "rank": str(r["rank"]["tail"]),
                         "new_score": str(n["score"]["tail"]), "new_rank": str(n["rank"]["tail"])})
        out.append({"triple_to_explain": pred, "conversions": conv})
    return out, new_model
