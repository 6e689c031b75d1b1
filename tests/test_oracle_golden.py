"""Pin the oracle: replay the seeds of tests/golden/*.npz (outputs of the UNMODIFIED
reference, produced by tests/golden/make_golden.py) through oracle/kelpie_oracle.py."""
import numpy as np
import pytest
import torch

from oracle import kelpie_oracle as ko
from tests.golden_util import load, seed_all, trace_of

KINDS = ["TransE", "ComplEx", "ConvE"]
# same torch ops in the same order: agreement is at rounding level, far inside the
# 1e-4 relative tolerance the CUDA path is held to.
RTOL = 2e-5


def _check_trace(engine, z, tag):
    ref = trace_of(z, tag)
    assert len(engine.trace) == len(ref)
    for (_, init, final, res), (r_init, r_final, r_res) in zip(engine.trace, ref):
        np.testing.assert_allclose(init.numpy(), r_init, rtol=0, atol=0)
        scale = np.abs(r_final).max()
        assert np.abs(final.numpy() - r_final.reshape(-1)).max() <= RTOL * scale
        assert res["target_rank"] == int(r_res[1])
        assert abs(res["target_score"] - r_res[0]) <= RTOL * max(1.0, abs(r_res[0]))
        assert abs(res["best_score"] - r_res[2]) <= RTOL * max(1.0, abs(r_res[2]))


@pytest.mark.parametrize("kind", KINDS)
def test_engines_match_reference(kind):
    z, meta, kg, w, order = load(kind)
    for case in meta["cases"]:
        eng = ko.Engine(w, kg, meta["hp"], mode=case["mode"], fact_order=order)
        seed_all(case["seed"])
        pred = tuple(case["pred"])
        if case["mode"] == "sufficient":
            eng.entities_to_convert = ko.select_entities_to_convert(w, kg, pred, 3, 200)
            assert eng.entities_to_convert == case["entities_to_convert"]
        rels = [eng.compute_relevance(pred, [tuple(t) for t in r]) for r in case["rules"]]
        _check_trace(eng, z, case["tag"])
        ref = z[case["tag"] + "relevance"]
        np.testing.assert_allclose(rels, ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("kind", KINDS)
def test_scores_and_ranks_match_reference(kind):
    z, meta, kg, w, _ = load(kind)
    with torch.no_grad():
        sc = ko.all_scores(w, w.ent, z["all_scores_q"]).numpy()
    np.testing.assert_allclose(sc, z["all_scores"], rtol=1e-5, atol=1e-6)
    res = ko.predict_triples(w, kg, z["predict_q"])
    ranks = np.array([[r["rank"]["tail"], r["rank"]["head"]] for r in res])
    scores = np.array([[r["score"]["tail"], r["score"]["head"]] for r in res])
    np.testing.assert_array_equal(ranks, z["predict_ranks"])
    np.testing.assert_allclose(scores, z["predict_scores"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("kind", KINDS)
def test_convertible_entities_match_reference(kind):
    z, meta, kg, w, _ = load(kind)
    got = sorted(ko.convertible_entities(w, kg, tuple(meta["convertible_pred"]), 200))
    np.testing.assert_array_equal(got, z["convertible"])


def test_n2_regulariser_matches_reference():
    """regularizers.py:25-34 (N2, weight 0.05) through the unmodified reference vs the oracle: tests/golden/make_golden_n2.py."""
    import json
    import os
    from tests.golden_util import GOLDEN
    z, meta, kg, w, order = load("ComplEx")
    g = np.load(os.path.join(GOLDEN, "complex_n2_small.npz"))
    gm = json.loads(bytes(g["meta"]).decode())
    eng = ko.Engine(w, kg, gm["hp"], mode="necessary", fact_order=order)
    seed_all(gm["seed"])
    rels = [eng.compute_relevance(tuple(gm["pred"]), [tuple(t) for t in r]) for r in gm["rules"]]
    _check_trace(eng, g, "n2_")
    np.testing.assert_allclose(rels, g["n2_relevance"], rtol=1e-5, atol=1e-5)
    # the regulariser did change the result: the N3 run of the same case gives other rows
    assert np.abs(g["n2_1_final"] - z["nec0_1_final"]).max() > 1e-5
