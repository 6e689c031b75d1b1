"""verify_explanations compute flow (kelpie_b200/verify_explanations.py) against golden outputs of the
unmodified reference classes (tests/golden/make_golden_verify.py): necessary mode (remove the best rules,
retrain TransE from scratch on the device, re-rank) and sufficient mode (add the rule to the conversion
entities, drop existing objects of *-to-one relations, retrain, re-rank the conversions).

Ranks before the retrain are bit-exact.  After 57 dependent Adam steps the retrained tables agree within
the 2e-4 stated in tests/test_fit_transe.py, so the new scores are compared at 1e-3 relative and the new
ranks may differ by at most the number of entities whose new score lies within that band of the target's."""
import json
import os

import numpy as np
import pytest
import torch

from tests.golden_util import GOLDEN

pytestmark = pytest.mark.gpu


def _setup():
    from kelpie_b200.data import Dataset
    from kelpie_b200.link_prediction import MODEL_REGISTRY
    z = np.load(os.path.join(GOLDEN, "verify_small.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    ds = Dataset("golden-verify", z["train"], z["valid"], z["test"], int(z["n_ent"]), int(z["n_rel"]))
    cls = MODEL_REGISTRY["TransE"]["class"]
    hp = cls.get_hyperparams_class()(**meta["config"]["model_params"])

    def make(dataset, ent, rel, burn=False):
        if burn:  # the reference's TransE(init_random=True) draws on the (CPU-patched) host generator before training
            n, r2, d = dataset.num_entities, 2 * dataset.num_relations, hp.dimension
            torch.rand(n, d), torch.rand(r2, d)
            torch.nn.init.xavier_normal_(torch.empty(n, d)), torch.nn.init.xavier_normal_(torch.empty(r2, d))
        m = cls(dataset, hp, init_random=False)
        with torch.no_grad():
            m.entity_embeddings.copy_(torch.from_numpy(ent))
            m.relation_embeddings.copy_(torch.from_numpy(rel))
        m.eval()
        return m

    return z, meta, ds, make


def _band(new_model, triple, score, rel_tol):
    sc = new_model.all_scores(np.array([triple]))[0].cpu().numpy()
    return int((np.abs(sc - score) <= rel_tol * max(1.0, abs(score))).sum()) - 1


def _compare(evals_flat, scores, ranks, new_model, triples):
    for ev, sc, rk, t in zip(evals_flat, scores, ranks, triples):
        assert int(ev["rank"]) == int(rk[0])                                  # before: bit-exact
        assert abs(float(ev["score"]) - sc[0]) <= 1e-4 * max(1.0, abs(sc[0]))
        assert abs(float(ev["new_score"]) - sc[1]) <= 1e-3 * max(1.0, abs(sc[1]))
        assert abs(int(ev["new_rank"]) - int(rk[1])) <= _band(new_model, t, float(ev["new_score"]), 1e-3)


def test_verify_necessary_matches_reference():
    from kelpie_b200 import verify_explanations as ve
    from tests.golden_util import seed_all
    z, meta, ds, make = _setup()
    model = make(ds, z["w_ent"], z["w_rel"])
    c = meta["necessary"]
    preds = [tuple(p) for p in c["preds"]]
    rules = {p: [tuple(t) for t in c["rules"][str(list(p))]] for p in preds}
    seed_all(c["seed"])
    evals, new_model = ve.verify_necessary(model, ds, rules, meta["config"],
                                           model_factory=lambda d: make(d, z["nec_ent0"], z["nec_rel0"], burn=True))
    assert [e["triple_to_explain"] for e in evals] == preds
    ent = new_model.entity_embeddings.detach().cpu().numpy()
    assert np.abs(ent - z["nec_ent"]).max() <= 2e-4 * np.abs(z["nec_ent"]).max()
    _compare(evals, z["nec_scores"], z["nec_ranks"], new_model, preds)
    assert len(ds.training_triples) == len(z["train"])  # the caller's dataset is untouched


def test_verify_sufficient_matches_reference():
    from kelpie_b200 import verify_explanations as ve
    from tests.golden_util import seed_all
    z, meta, ds, make = _setup()
    model = make(ds, z["w_ent"], z["w_rel"])
    c = meta["sufficient"]
    preds = [tuple(p) for p in c["preds"]]
    rules = {p: [tuple(t) for t in c["rules"][str(list(p))]] for p in preds}
    ents = {p: c["entities"][str(list(p))] for p in preds}
    seed_all(c["seed"])
    evals, new_model = ve.verify_sufficient(model, ds, rules, ents, meta["config"],
                                            model_factory=lambda d: make(d, z["suf_ent0"], z["suf_rel0"], burn=True))
    assert len(new_model.dataset.training_triples) == int(z["suf_n_train"])
    flat = [cv for e in evals for cv in e["conversions"]]
    _compare(flat, z["suf_scores"], z["suf_ranks"], new_model, [tuple(t) for t in c["to_convert"]])


@pytest.mark.parametrize("kind", ["ComplEx", "ConvE"])
def test_verify_necessary_flow_other_models(kind):
    """The same flow for the other two model kinds (their device trainers: tests/test_fit_complex.py, test_fit_conve.py):
    the ranks / scores BEFORE the retrain are those of predict_triples, removing the rule and retraining changes the
    model, and the caller's dataset and model stay untouched."""
    from kelpie_b200 import verify_explanations as ve
    from kelpie_b200.data import Dataset
    from kelpie_b200.link_prediction import MODEL_REGISTRY
    from tests.golden_util import load, seed_all
    g, meta, kg, w, _ = load(kind)
    ds = Dataset("golden", g["train"], g["valid"], g["test"], kg.num_entities, kg.num_relations)
    cls = MODEL_REGISTRY[kind]["class"]
    seed_all(3)
    model = cls(ds, cls.get_hyperparams_class()(**meta["params"]), init_random=True)
    model.eval()
    training = (dict(optimizer_name="Adagrad", batch_size=128, epochs=2, lr=0.043, decay1=0.9, decay2=0.999, regularizer_name="N3",
                     regularizer_weight=0) if kind == "ComplEx"
                else dict(batch_size=128, label_smoothing=0.1, lr=0.003, decay=0.995, epochs=2))
    preds = [tuple(int(x) for x in t) for t in g["test"][:4]]
    rules = {p: [tuple(int(x) for x in t) for t in ds.entity_to_training_triples[p[0]][:2]] for p in preds}
    before = model.predict_triples(np.array(preds))
    ent0 = model.entity_embeddings.detach().clone()
    evals, new_model = ve.verify_necessary(model, ds, rules, {"model_params": meta["params"], "training": training})
    for ev, b, p in zip(evals, before, preds):
        assert ev["triple_to_explain"] == p and ev["rank"] == str(b["rank"]["tail"]) and ev["score"] == str(b["score"]["tail"])
        assert float(ev["new_score"]) == float(ev["new_score"])  # finite
    removed = sum(len(set(r)) for r in rules.values())
    assert len(ds.training_triples) == len(g["train"]) and len(new_model.dataset.training_triples) <= len(g["train"]) - 1
    assert torch.equal(ent0, model.entity_embeddings.detach()) and new_model is not model
    assert removed >= 1 and any(ev["new_rank"] != ev["rank"] or ev["new_score"] != ev["score"] for ev in evals)
