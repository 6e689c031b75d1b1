"""Golden relevances of the data-poisoning baseline from the UNMODIFIED reference (CPU-patched):
NecessaryDPEngine / SufficientDPEngine (src/relevance_engines/data_poisoning_engine.py) on the ComplEx model and KG
of complex_small.npz (make_golden.py).  For 12 test predictions: every training fact of the head (perspective "head")
and of the tail (perspective "tail") in necessary mode; the head's facts in sufficient mode with 4 conversion entities
(which exercises the reference's reassignment quirk in SufficientDPEngine.compute_relevance).

    python tests/golden/make_golden_dp.py
"""
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
from src.relevance_engines import NecessaryDPEngine, SufficientDPEngine  # noqa: E402

EPSILON = 0.05

if __name__ == "__main__":
    z = np.load(os.path.join(HERE, "complex_small.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    n_ent, n_rel = int(z["n_ent"]), int(z["n_rel"])
    refshim.register_dataset("golden-dp", z["train"], z["valid"], z["test"], n_ent, n_rel)
    ds = Dataset("golden-dp")
    model = ComplEx(ds, ComplExHyperParams(**meta["params"]))
    with torch.no_grad():
        model.entity_embeddings.copy_(torch.from_numpy(z["w_ent"]))
        model.relation_embeddings.copy_(torch.from_numpy(z["w_rel"]))
    model.eval()
    nec, suf = NecessaryDPEngine(model, ds, EPSILON), SufficientDPEngine(model, ds, EPSILON)
    rows = []  # (mode, pred s p o, perspective 0 head / 1 tail, fact s p o, relevance); sufficient: + the 4 entities
    rng = np.random.default_rng(5)
    preds = [tuple(int(x) for x in t) for t in z["test"][:12]]
    ents = []
    for pred in preds:
        s, p, o = pred
        for persp, e in (("head", s), ("tail", o)):
            for fact in ds.entity_to_training_triples[e][:8]:
                rows.append((0,) + pred + (0 if persp == "head" else 1,) + tuple(int(x) for x in fact) + (float(nec.compute_relevance(pred, persp, fact)),))
        conv = [int(x) for x in rng.choice([e for e in range(n_ent) if e != s], 4, replace=False)]
        ents.append(conv)
        suf.entities_to_convert = conv
        for fact in ds.entity_to_training_triples[s][:8]:
            rows.append((1,) + pred + (0,) + tuple(int(x) for x in fact) + (float(suf.compute_relevance(pred, "head", fact)),))
    rows = np.array(rows, dtype=np.float64)
    path = os.path.join(HERE, "dp_small.npz")
    np.savez_compressed(path, rows=rows, entities=np.array(ents, dtype=np.int64), preds=np.array(preds, dtype=np.int64), epsilon=np.float64(EPSILON))
    print(len(rows), "relevances ->", path, os.path.getsize(path), "bytes; |rel| range", np.abs(rows[:, -1]).min(), np.abs(rows[:, -1]).max())
