"""Host-only profile of the explain path: the engine and the builder driven with a stub device context on the CPU
(post-training returns the init rows, the rank pass zeros), so that what is timed is exactly the Python / numpy side of
tools/bench_explain.py: KelpieDataset edits, plan drawing in the reference's RNG order, staging, result boxing, the builder."""
import sys, os, time, random, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kelpie_b200.data import Dataset
from kelpie_b200.link_prediction import MODEL_REGISTRY
from kelpie_b200.relevance_engines import NecessaryPostTrainingEngine
from kelpie_b200.explanation_builders import StochasticBuilder

ap = argparse.ArgumentParser()
ap.add_argument("--preds", type=int, default=100)
ap.add_argument("--model", default="TransE")
ap.add_argument("--profile", action="store_true")
a = ap.parse_args()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ds = Dataset.from_npz(os.path.join(ROOT, "tests", "golden", "dbpedia50_ids.npz"), name="DBpedia50")
CFG = {"TransE": (dict(dimension=256, norm=2), dict(batch_size=2048, epochs=65, lr=0.01, margin=5, negative_triples_ratio=5, regularizer_weight=1.0)),
       "ComplEx": (dict(dimension=200, init_scale=1e-3), dict(optimizer_name="Adagrad", batch_size=512, epochs=43, lr=0.043, decay1=0.9, decay2=0.999,
                                                              regularizer_name="N3", regularizer_weight=0))}


class StubContext:
    def __init__(self, D):
        self.D = D

    def post_train(self, hp, dropout_seed=0, **arrs):
        return torch.zeros((len(arrs["row_off"]) - 1 if "row_off" in arrs else 1, self.D))

    def filtered_rank(self, triples, mode, mimic_rows=None, flt_off=None, flt_ids=None):
        n = len(triples)
        return torch.zeros(n), torch.zeros(n), torch.ones(n, dtype=torch.int64)


preds = []
for s, p, o in ds.testing_triples:
    if 3 <= len(ds.entity_to_training_triples[s]) <= 20:
        preds.append((int(s), int(p), int(o)))
    if len(preds) == a.preds:
        break
params, hp = CFG[a.model]
cls = MODEL_REGISTRY[a.model]["class"]
torch.manual_seed(7)
m = cls(ds, cls.get_hyperparams_class()(**params), init_random=False)
m.eval()
object.__setattr__(m, "_kp_ctx", StubContext(m.dimension if a.model != "ComplEx" else 2 * m.dimension))
eng = NecessaryPostTrainingEngine(m, ds, hp)
eng.rng_device = "cpu"
builder = StochasticBuilder(xsi=0.4, engine=eng, max_explanation_length=4)  # every relevance is 0.5: one singleton batch per prediction
if a.profile:
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
best = None
for rep in range(10):
    torch.manual_seed(1); np.random.seed(1); random.seed(1)
    n_rel, t0 = 0, time.perf_counter()
    for pred in preds:
        eng.set_cache()
        n_rel += builder.build_explanations(pred, ds.entity_to_training_triples[pred[0]])["#relevances"]
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
if a.profile:
    pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(28)
print(f"{a.model}: {n_rel} relevances, {1e6 * best / n_rel:.0f} us of host time per relevance, {1e3 * best / len(preds):.2f} ms per prediction")
