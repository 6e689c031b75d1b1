"""Data-poisoning baseline (data_poisoning_engine.py) against relevances produced by the unmodified
reference's NecessaryDPEngine / SufficientDPEngine on the ComplEx fixture (tests/golden/make_golden_dp.py).
A relevance is a DIFFERENCE of two nearly equal fp32 scores (|score| up to ~0.3, relevances 2e-5 .. 1.6e-2), so the
stated tolerance is absolute: 2e-7 (a few ulps of the scores) + 1e-4 relative."""
import os

import numpy as np
import pytest
import torch

from tests.golden_util import GOLDEN, load


def _rows():
    z = np.load(os.path.join(GOLDEN, "dp_small.npz"))
    return z, z["rows"]


def _close(got, want):
    return abs(got - want) <= 2e-7 + 1e-4 * abs(want)


def test_oracle_matches_reference():
    from oracle import kelpie_oracle as ko
    z, rows = _rows()
    _, _, _, w, _ = load("ComplEx")
    eps = float(z["epsilon"])
    preds = [tuple(p) for p in z["preds"]]
    for r in rows:
        mode, pred, persp, fact, want = int(r[0]), tuple(int(x) for x in r[1:4]), int(r[4]), tuple(int(x) for x in r[5:8]), r[8]
        if mode == 0:
            got = ko.dp_relevance(w.ent, w.rel, pred, fact, pred[0] if persp == 0 else pred[2], eps)
        else:  # the reference's loop converts pred / fact once (first entity) and then re-evaluates the same pair
            e0 = int(z["entities"][preds.index(pred)][0])
            swap = lambda t: tuple(e0 if x == pred[0] and i != 1 else x for i, x in enumerate(t))
            got = ko.dp_relevance(w.ent, w.rel, swap(pred), swap(fact), e0, eps, sufficient=True)
        assert _close(got, want), (r, got)


@pytest.mark.gpu
def test_cuda_engines_match_reference():
    from kelpie_b200.data import Dataset
    from kelpie_b200.link_prediction import MODEL_REGISTRY
    from kelpie_b200.relevance_engines import NecessaryDPEngine, SufficientDPEngine
    z, rows = _rows()
    g, meta, kg, w, _ = load("ComplEx")
    ds = Dataset("golden-dp", g["train"], g["valid"], g["test"], kg.num_entities, kg.num_relations)
    cls = MODEL_REGISTRY["ComplEx"]["class"]
    m = cls(ds, cls.get_hyperparams_class()(**meta["params"]), init_random=False)
    with torch.no_grad():
        m.entity_embeddings.copy_(w.ent)
        m.relation_embeddings.copy_(w.rel)
    m.eval()
    eps = float(z["epsilon"])
    nec, suf = NecessaryDPEngine(m, ds, eps), SufficientDPEngine(m, ds, eps)
    preds = [tuple(int(x) for x in p) for p in z["preds"]]
    for pred in preds:
        for mode, persp in ((0, 0), (0, 1), (1, 0)):
            sel = [r for r in rows if int(r[0]) == mode and tuple(int(x) for x in r[1:4]) == pred and int(r[4]) == persp]
            facts = [tuple(int(x) for x in r[5:8]) for r in sel]
            if mode == 0:
                got = nec.compute_relevances(pred, "head" if persp == 0 else "tail", facts)
                one = nec.compute_relevance(pred, "head" if persp == 0 else "tail", facts[0])
            else:
                suf.entities_to_convert = [int(x) for x in z["entities"][preds.index(pred)]]
                got = suf.compute_relevances(pred, "head", facts)
                one = suf.compute_relevance(pred, "head", facts[0])
            assert len(got) == len(sel) and one == got[0]
            for r, x in zip(sel, got):
                assert _close(float(x), r[8]), (r, x)


@pytest.mark.gpu
def test_other_models_are_refused_like_the_reference():
    """TransE / ConvE define score_embs, not score_embeddings: the reference raises AttributeError inside get_gradient."""
    from kelpie_b200.data import Dataset
    from kelpie_b200.link_prediction import MODEL_REGISTRY
    from kelpie_b200.relevance_engines import NecessaryDPEngine
    g, meta, kg, w, _ = load("TransE")
    ds = Dataset("golden-dp", g["train"], g["valid"], g["test"], kg.num_entities, kg.num_relations)
    cls = MODEL_REGISTRY["TransE"]["class"]
    m = cls(ds, cls.get_hyperparams_class()(**meta["params"]), init_random=True)
    with pytest.raises(RuntimeError, match="score_embeddings"):
        NecessaryDPEngine(m, ds, 0.1).compute_relevance((0, 0, 1), "head", (0, 0, 2))
