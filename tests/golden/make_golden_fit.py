"""Golden vectors of FULL-model TransE training (verify_explanations' retrain) from the
UNMODIFIED reference (CPU-patched): PairwiseRankingOptimizer(model, hp).train(train) for 3 epochs on the
300-entity synthetic KG of make_golden.py (dimension 64, batch 256 so that an epoch is 19 dependent steps),
L2 and L1 norm.  Stores the initial and the trained tables.

    python tests/golden/make_golden_fit.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refshim  # noqa: E402

refshim.install(cpu=True)

from src.data import Dataset  # noqa: E402
from src.link_prediction.models import TransE  # noqa: E402
from src.link_prediction.models.transe import TransEHyperParams  # noqa: E402
from src.link_prediction.optimization import PairwiseRankingOptimizer  # noqa: E402
from src.link_prediction.optimization.pairwise_ranking_optimizer import PairwiseRankingOptimizerHyperParams  # noqa: E402

from tests.golden.make_golden import seed_all, synthetic_kg  # noqa: E402

CX_HP = dict(optimizer_name="Adagrad", batch_size=128, epochs=3, lr=0.043, decay1=0.9, decay2=0.999, regularizer_name="N3",
             regularizer_weight=0)

HP = dict(batch_size=256, epochs=3, lr=0.01, margin=5, negative_triples_ratio=5, regularizer_weight=1.0)

if __name__ == "__main__":
    n_ent, n_rel = 300, 10
    train, valid, test = synthetic_kg(7, n_ent, n_rel, 2400, 150, 150)
    refshim.register_dataset("golden-fit", train, valid, test, n_ent, n_rel)
    ds = Dataset("golden-fit")
    out = dict(train=train, valid=valid, test=test, n_ent=np.int64(n_ent), n_rel=np.int64(n_rel))
    for norm in (2, 1):
        seed_all(5 + norm)
        model = TransE(ds, TransEHyperParams(dimension=64, norm=norm), init_random=True)
        out[f"n{norm}_ent0"] = model.entity_embeddings.detach().numpy().copy()
        out[f"n{norm}_rel0"] = model.relation_embeddings.detach().numpy().copy()
        seed_all(50 + norm)
        opt = PairwiseRankingOptimizer(model=model, hp=PairwiseRankingOptimizerHyperParams(**HP), verbose=False)
        opt.train(training_triples=ds.training_triples)
        out[f"n{norm}_ent"] = model.entity_embeddings.detach().numpy().copy()
        out[f"n{norm}_rel"] = model.relation_embeddings.detach().numpy().copy()
        print("norm", norm, "moved", np.abs(out[f"n{norm}_ent"] - out[f"n{norm}_ent0"]).max())
    path = os.path.join(HERE, "transe_fit_small.npz")
    np.savez_compressed(path, **out)
    print("->", path, os.path.getsize(path), "bytes")

    # ---- ComplEx: MultiClassNLLOptimizer.train (Adagrad as in the shipped configs, and Adam)
    from src.link_prediction.models import ComplEx
    from src.link_prediction.models.complex import ComplExHyperParams
    from src.link_prediction.optimization import MultiClassNLLOptimizer
    from src.link_prediction.optimization.multiclass_nll_optimizer import MultiClassNLLOptimizerHyperParams
    out = dict(train=train, valid=valid, test=test, n_ent=np.int64(n_ent), n_rel=np.int64(n_rel))
    for name in ("Adagrad", "Adam"):
        seed_all(9)
        model = ComplEx(ds, ComplExHyperParams(dimension=32, init_scale=1e-3), init_random=True)
        with torch.no_grad():  # the reference's init (scale 1e-3) barely moves in 3 epochs: use a trained-like scale
            model.entity_embeddings.mul_(300.0)
            model.relation_embeddings.mul_(300.0)
        out[f"{name}_ent0"] = model.entity_embeddings.detach().numpy().copy()
        out[f"{name}_rel0"] = model.relation_embeddings.detach().numpy().copy()
        seed_all(60)
        hp = dict(CX_HP, optimizer_name=name, lr=0.043 if name == "Adagrad" else 0.01)
        opt = MultiClassNLLOptimizer(model=model, hp=MultiClassNLLOptimizerHyperParams(**hp), verbose=False)
        opt.train(training_triples=ds.training_triples)
        out[f"{name}_ent"] = model.entity_embeddings.detach().numpy().copy()
        out[f"{name}_rel"] = model.relation_embeddings.detach().numpy().copy()
        print(name, "moved", np.abs(out[f"{name}_ent"] - out[f"{name}_ent0"]).max(), "of", np.abs(out[f"{name}_ent0"]).max())
    path = os.path.join(HERE, "complex_fit_small.npz")
    np.savez_compressed(path, **out)
    print("->", path, os.path.getsize(path), "bytes")
