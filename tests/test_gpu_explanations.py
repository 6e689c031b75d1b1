"""End to end: the explanations selected with the batched CUDA engine are identical to those the
sequential CPU oracle engine selects (same seeds, same candidate facts), necessary + sufficient."""
import numpy as np
import pytest

from oracle import kelpie_oracle as ko
from tests.golden_util import load, seed_all
from tests.test_gpu_parity import _dataset, _model

pytestmark = pytest.mark.gpu


class _Labels:
    def labels_triple(self, t):
        return tuple(int(x) for x in t)

    def labels_triples(self, ts):
        return [self.labels_triple(t) for t in ts]


@pytest.mark.parametrize("kind", ["TransE", "ComplEx", "ConvE"])
@pytest.mark.parametrize("mode", ["necessary", "sufficient"])
def test_selected_explanations_identical(kind, mode):
    from kelpie_b200.explanation_builders import StochasticBuilder
    from kelpie_b200.relevance_engines import NecessaryPostTrainingEngine, SufficientPostTrainingEngine
    z, meta, kg, w, order = load(kind)
    hp = dict(meta["hp"], epochs=4)
    case = meta["cases"][0 if mode == "necessary" else 2]
    pred = tuple(case["pred"])
    facts = order[pred[0]][:5]
    xsi = 5.0 if mode == "necessary" else 0.9

    # sequential oracle
    oracle = ko.Engine(w, kg, hp, mode=mode, fact_order=order)
    oracle.dataset = _Labels()
    seed_all(77)
    if mode == "sufficient":
        oracle.entities_to_convert = ko.select_entities_to_convert(w, kg, pred, 2, 200)
    want = StochasticBuilder(xsi, oracle).build_explanations(pred, list(facts))

    # batched CUDA engine
    ds = _dataset(z)
    for e, f in order.items():
        ds.entity_to_training_triples[e] = [tuple(t) for t in f]
    ds.labels_triple = _Labels().labels_triple
    ds.labels_triples = _Labels().labels_triples
    model = _model(kind, z, meta, ds)
    cls = NecessaryPostTrainingEngine if mode == "necessary" else SufficientPostTrainingEngine
    eng = cls(model, ds, hp)
    eng.rng_device = "cpu"
    seed_all(77)
    if mode == "sufficient":
        eng.select_entities_to_convert(pred, 2, 200)
        assert [int(e) for e in eng.entities_to_convert] == [int(e) for e in oracle.entities_to_convert]
    got = StochasticBuilder(xsi, eng, batch_size=6).build_explanations(pred, list(facts))

    assert got["#relevances"] == want["#relevances"]
    assert [r for r, _ in got["rule_to_relevance"]] == [r for r, _ in want["rule_to_relevance"]]
    np.testing.assert_allclose([v for _, v in got["rule_to_relevance"]], [v for _, v in want["rule_to_relevance"]],
                               rtol=1e-4, atol=1e-4)
