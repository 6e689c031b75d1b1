"""Golden vectors for the N2 regulariser (regularizers.py:25-34, selected by `regularizer_name` at
multiclass_nll_optimizer.py:46-49) produced by the UNMODIFIED reference in the build container:

    python tests/golden/make_golden_n2.py

Same 300-entity KG and ComplEx weights as complex_small.npz (read from it); the first necessary-mode case is
re-run with regularizer_name = "N2", regularizer_weight = 0.05.  Stored: the hyper-parameters, the post-training
trace (initial / final mimic rows, (score, rank, best score)) and the relevances."""
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
from src.link_prediction.models import ComplEx  # noqa: E402
from src.link_prediction.models.complex import ComplExHyperParams  # noqa: E402
from src.relevance_engines import NecessaryPostTrainingEngine  # noqa: E402

from tests.golden.make_golden import pack_trace, seed_all, traced  # noqa: E402

if __name__ == "__main__":
    z = np.load(os.path.join(HERE, "complex_small.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    refshim.register_dataset("golden-n2", z["train"], z["valid"], z["test"], int(z["n_ent"]), int(z["n_rel"]))
    ds = Dataset("golden-n2")
    model = ComplEx(ds, ComplExHyperParams(**meta["params"]))
    with torch.no_grad():
        model.entity_embeddings.copy_(torch.from_numpy(z["w_ent"]))
        model.relation_embeddings.copy_(torch.from_numpy(z["w_rel"]))
    model.eval()
    hp = dict(meta["hp"], regularizer_name="N2", regularizer_weight=0.05)
    case = meta["cases"][0]
    pred = tuple(case["pred"])
    for e, facts in meta["fact_order"].items():  # the fact order of the first run (Python set order is not stable)
        ds.entity_to_training_triples[int(e)] = [tuple(t) for t in facts]
    eng = NecessaryPostTrainingEngine(model, ds, hp)
    seed_all(case["seed"])
    eng.set_cache()
    trace = traced(eng)
    rules = [[tuple(t) for t in r] for r in case["rules"]]
    rels = [eng.compute_relevance(pred, r) for r in rules]
    out = {}
    pack_trace(trace, "n2_", out)
    out["n2_relevance"] = np.array(rels, dtype=np.float64)
    out["meta"] = np.frombuffer(json.dumps(dict(hp=hp, pred=pred, rules=rules, seed=case["seed"])).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "complex_n2_small.npz")
    np.savez_compressed(path, **out)
    print("N2 ->", path, os.path.getsize(path), "bytes", np.round(rels, 4))
