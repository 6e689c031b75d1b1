"""GPU parity: the CUDA path (through the C ABI) against the reference's golden vectors and
against the oracle on seeded inputs.  Tolerances: integer ranks bit-exact; scores and
post-trained mimic rows within 1e-4 relative (north_star), stated per assertion."""
import numpy as np
import pytest
import torch

from tests.golden_util import load, seed_all, trace_of

pytestmark = pytest.mark.gpu

RTOL = 1e-4
READY = {"TransE": True, "ComplEx": True, "ConvE": True}


def _dataset(z):
    from kelpie_b200.data import Dataset
    return Dataset("golden", z["train"], z["valid"], z["test"], int(z["n_ent"]), int(z["n_rel"]))


def _model(kind, z, meta, ds):
    from kelpie_b200.link_prediction import MODEL_REGISTRY
    cls = MODEL_REGISTRY[kind]["class"]
    m = cls(ds, cls.get_hyperparams_class()(**meta["params"]), init_random=False)
    with torch.no_grad():
        m.entity_embeddings.copy_(torch.from_numpy(z["w_ent"]))
        m.relation_embeddings.copy_(torch.from_numpy(z["w_rel"]))
        if kind == "ConvE":
            m.convolutional_layer.weight.copy_(torch.from_numpy(z["w_conv_w"]))
            m.convolutional_layer.bias.copy_(torch.from_numpy(z["w_conv_b"]))
            m.hidden_layer.weight.copy_(torch.from_numpy(z["w_fc_w"]))
            m.hidden_layer.bias.copy_(torch.from_numpy(z["w_fc_b"]))
            for i, bn in enumerate((m.batch_norm_1, m.batch_norm_2, m.batch_norm_3), 1):
                bn.weight.copy_(torch.from_numpy(z[f"w_bn{i}_w"]))
                bn.bias.copy_(torch.from_numpy(z[f"w_bn{i}_b"]))
                bn.running_mean.copy_(torch.from_numpy(z[f"w_bn{i}_mean"]))
                bn.running_var.copy_(torch.from_numpy(z[f"w_bn{i}_var"]))
    m.eval()
    return m


def _close(a, b, rtol=RTOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    assert np.abs(a - b).max() <= rtol * scale, (np.abs(a - b).max(), scale)


@pytest.mark.parametrize("kind", ["TransE", "ComplEx", "ConvE"])
def test_all_scores_and_predict_triples_golden(kind):
    z, meta, kg, w, order = load(kind)
    ds = _dataset(z)
    m = _model(kind, z, meta, ds)
    sc = m.all_scores(z["all_scores_q"]).cpu().numpy()
    _close(sc, z["all_scores"])
    res = m.predict_triples(z["predict_q"])
    ranks = np.array([[r["rank"]["tail"], r["rank"]["head"]] for r in res])
    scores = np.array([[r["score"]["tail"], r["score"]["head"]] for r in res])
    np.testing.assert_array_equal(ranks, z["predict_ranks"])  # bit-exact integer ranks
    _close(scores, z["predict_scores"])


@pytest.mark.parametrize("kind", ["TransE", "ComplEx", "ConvE"])
def test_convertible_entities_golden(kind):
    from kelpie_b200.relevance_engines import RelevanceEngine
    z, meta, kg, w, order = load(kind)
    ds = _dataset(z)
    m = _model(kind, z, meta, ds)
    got = sorted(RelevanceEngine(m, ds).convertible_entities(tuple(meta["convertible_pred"]), 200))
    np.testing.assert_array_equal(got, z["convertible"])


def _run_engine_case(kind, case, z, meta, ds, m, order):
    from kelpie_b200.relevance_engines import NecessaryPostTrainingEngine, SufficientPostTrainingEngine
    cls = NecessaryPostTrainingEngine if case["mode"] == "necessary" else SufficientPostTrainingEngine
    eng = cls(m, ds, meta["hp"])
    eng.rng_device = "cpu"  # the golden run drew KelpieTransE's xavier row on the CPU generator
    for e, facts in order.items():  # fix the fact order the reference's Python sets produced
        ds.entity_to_training_triples[e] = [tuple(t) for t in facts]
    pred = tuple(case["pred"])
    seed_all(case["seed"])
    eng.set_cache()
    if case["mode"] == "sufficient":
        eng.select_entities_to_convert(pred, 3, 200)
        assert [int(e) for e in eng.entities_to_convert] == case["entities_to_convert"]
    rules = [[tuple(t) for t in r] for r in case["rules"]]
    return eng, rules, pred


@pytest.mark.parametrize("kind", ["TransE", "ComplEx", "ConvE"])
@pytest.mark.parametrize("batched", [False, True])
def test_engine_relevance_golden(kind, batched):
    if not READY[kind]:
        pytest.skip("post-training kernel not built yet")
    z, meta, kg, w, order = load(kind)
    for case in meta["cases"]:
        ds = _dataset(z)
        m = _model(kind, z, meta, ds)
        eng, rules, pred = _run_engine_case(kind, case, z, meta, ds, m, order)
        if batched:
            rels = eng.compute_relevances(pred, rules)
        else:
            rels = [eng.compute_relevance(pred, r) for r in rules]
        ref = z[case["tag"] + "relevance"]
        # relevance = rank delta (exact integer) + sigmoid(score delta) / (rank for sufficient)
        np.testing.assert_allclose(rels, ref, rtol=RTOL, atol=RTOL)


@pytest.mark.parametrize("kind", ["TransE", "ComplEx", "ConvE"])
def test_post_trained_rows_golden(kind):
    """Mimic rows after post-training and the (score, rank) of the target, job by job."""
    if not READY[kind]:
        pytest.skip("post-training kernel not built yet")
    z, meta, kg, w, order = load(kind)
    case = meta["cases"][0]
    ds = _dataset(z)
    m = _model(kind, z, meta, ds)
    eng, rules, pred = _run_engine_case(kind, case, z, meta, ds, m, order)
    ref = trace_of(z, case["tag"])
    # sequential calls: job order = base, pt(rule0), pt(rule1), ... exactly like the reference
    got_rows, got_res = [], []
    for r in rules:
        n_before = len(eng.base_pt_results)
        pt, base = eng.individual_results([(pred, r)])[0]
        rows = eng.last_rows.cpu().numpy()
        if len(eng.base_pt_results) > n_before:
            got_rows.append(rows[0]); got_res.append(base)
        got_rows.append(rows[-1]); got_res.append(pt)
    assert len(got_rows) == len(ref)
    for row, res, (r_init, r_final, r_res) in zip(got_rows, got_res, ref):
        _close(row, r_final.reshape(-1))
        assert int(res["target_rank"]) == int(r_res[1])
        assert abs(res["target_score"] - r_res[0]) <= RTOL * max(1.0, abs(r_res[0]))
        assert abs(float(res["best_score"]) - r_res[2]) <= RTOL * max(1.0, abs(r_res[2]))


@pytest.mark.parametrize("kind", ["TransE", "ComplEx", "ConvE"])
def test_evaluator_metrics_match_golden_ranks(kind):
    """Evaluator (evaluation.py:16-48,74-92) over predict_triples: metrics from the reference's ranks."""
    from kelpie_b200.link_prediction.evaluation import Evaluator
    z, meta, kg, w, order = load(kind)
    ds = _dataset(z)
    m = _model(kind, z, meta, ds)
    got = Evaluator(m).evaluate(z["predict_q"])
    ranks = np.concatenate([z["predict_ranks"][:, 1], z["predict_ranks"][:, 0]]).astype(float)
    assert got["mrr"] == pytest.approx(float(np.mean(1.0 / ranks)))
    assert got["mr"] == pytest.approx(float(np.mean(ranks)))
    assert got["h1"] == pytest.approx(float(np.mean(ranks <= 1)))
    assert got["h10"] == pytest.approx(float(np.mean(ranks <= 10)))


def test_transe_mimic_row_drawn_on_the_cuda_generator():
    """The DEFAULT rng_device ("cuda"): KelpieTransE overwrites its mimic row with xavier_normal_ on the row's device
    (transe.py:93-95: Parameter(init_tensor.cuda()) then xavier_normal_), i.e. on the CUDA generator when the reference
    runs on a GPU.  The engine's draw must be torch's own: same values, same generator state afterwards, two draws per
    compute_relevance (base model, then post-trained model) -- and it must end up in the post-training unchanged."""
    from torch.nn.init import xavier_normal_
    from kelpie_b200.relevance_engines import NecessaryPostTrainingEngine
    z, meta, kg, w, order = load("TransE")
    ds = _dataset(z)
    m = _model("TransE", z, meta, ds)
    eng = NecessaryPostTrainingEngine(m, ds, meta["hp"])
    assert eng.rng_device == "cuda"
    D = m.dimension
    torch.manual_seed(321)  # seeds the CPU and the CUDA generators
    init = torch.rand(1, D)
    got = [eng._init_row(init), eng._init_row(init)]
    tail = torch.rand(4, device="cuda")
    torch.manual_seed(321)
    init_ref = torch.rand(1, D)
    want = []
    for _ in range(2):
        p = torch.nn.Parameter(init_ref.cuda(), requires_grad=True)
        with torch.no_grad():
            xavier_normal_(p)
        want.append(p.detach())
    assert torch.equal(torch.rand(4, device="cuda"), tail)  # the CUDA generator advanced exactly as the reference's would
    for a, b in zip(got, want):
        assert a.is_cuda and torch.equal(a.view(-1), b.view(-1))
    # end to end with the default device: the rows the batch is post-trained from are those draws
    case = meta["cases"][0]
    pred, rules = tuple(case["pred"]), [[tuple(t) for t in r] for r in case["rules"]][:2]
    seed_all(77)  # torch (CPU + CUDA) and numpy: the corruptions come from torch.randint, the shuffles from np.random
    eng.set_cache()
    rels = eng.compute_relevances(pred, rules)
    seed_all(77)
    eng2 = NecessaryPostTrainingEngine(m, ds, meta["hp"])
    rels2 = [eng2.compute_relevance(pred, r) for r in rules]  # the reference's sequential calls
    np.testing.assert_allclose(rels, rels2, rtol=RTOL, atol=RTOL)
    assert all(np.isfinite(rels))


def test_complex_n2_regulariser_golden():
    """kp_hp.regularizer = N2 (regularizers.py:25-34): post-trained rows, ranks and relevances vs the unmodified reference."""
    import json
    import os
    from tests.golden_util import GOLDEN
    from kelpie_b200.relevance_engines import NecessaryPostTrainingEngine
    z, meta, kg, w, order = load("ComplEx")
    g = np.load(os.path.join(GOLDEN, "complex_n2_small.npz"))
    gm = json.loads(bytes(g["meta"]).decode())
    ds = _dataset(z)
    for e, facts in order.items():
        ds.entity_to_training_triples[e] = [tuple(t) for t in facts]
    m = _model("ComplEx", z, meta, ds)
    eng = NecessaryPostTrainingEngine(m, ds, gm["hp"])
    pred, rules = tuple(gm["pred"]), [[tuple(t) for t in r] for r in gm["rules"]]
    seed_all(gm["seed"])
    eng.set_cache()
    got_rows, got_res = [], []
    for r in rules:
        n_before = len(eng.base_pt_results)
        pt, base = eng.individual_results([(pred, r)])[0]
        rows = eng.last_rows.cpu().numpy()
        if len(eng.base_pt_results) > n_before:
            got_rows.append(rows[0]); got_res.append(base)
        got_rows.append(rows[-1]); got_res.append(pt)
    ref = trace_of(g, "n2_")
    assert len(got_rows) == len(ref)
    for row, res, (r_init, r_final, r_res) in zip(got_rows, got_res, ref):
        _close(row, r_final.reshape(-1))
        assert int(res["target_rank"]) == int(r_res[1])
        assert abs(res["target_score"] - r_res[0]) <= RTOL * max(1.0, abs(r_res[0]))
    seed_all(gm["seed"])
    eng.set_cache()
    np.testing.assert_allclose(eng.compute_relevances(pred, rules), g["n2_relevance"], rtol=RTOL, atol=RTOL)


@pytest.mark.parametrize("kind", ["TransE", "ComplEx", "ConvE"])
def test_model_score_and_forward(kind):
    """Model.score (one row of work per triple, kp_score_triples) and Model.forward against the golden all_scores
    of the unmodified reference (transe.py:38-46,67-75, complex.py:41-86, conve.py:65-75)."""
    z, meta, kg, w, order = load(kind)
    ds = _dataset(z)
    m = _model(kind, z, meta, ds)
    q = z["all_scores_q"]
    want = z["all_scores"][np.arange(len(q)), q[:, 2]]
    got = np.asarray(m.score(q)).reshape(-1)
    _close(got, want)
    out = m.forward(q)
    if kind == "ConvE":
        _close(out.cpu().numpy(), z["all_scores"])
    else:
        sc, factors = out
        assert len(factors) == 3 and all(f.shape[0] == len(q) for f in factors)
        if kind == "TransE":
            _close(sc.cpu().numpy(), want)
            np.testing.assert_array_equal(factors[0].detach().cpu().numpy(), z["w_ent"][q[:, 0]])
        else:
            _close(sc.cpu().numpy(), z["all_scores"])
            d = z["w_ent"].shape[1] // 2
            l = z["w_ent"][q[:, 0]]
            _close(factors[0].detach().cpu().numpy(), np.sqrt(l[:, :d] ** 2 + l[:, d:] ** 2))
