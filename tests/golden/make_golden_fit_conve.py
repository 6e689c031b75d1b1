"""Golden vectors of FULL-model ConvE training (verify_explanations' retrain) from the UNMODIFIED
reference (CPU-patched): BCEOptimizer(model, hp).train(train) (bce_optimizer.py:44-158) on the 300-entity synthetic KG
of make_golden.py.  dimension 60 (20 x 3 embedding image -> 32 filters x 38 x 1 = hidden 1216), all dropout rates 0
(the DBpedia50 config; parity with dropout > 0 would need torch's own masks), label smoothing 0.1, Adam + ExponentialLR.
Two runs from the same initial state:
  "a": batch 128, 3 epochs (train-mode batch-norm on every step);
  "b": batch = (number of pairs - 1), 1 epoch = TWO steps: all pairs but one (train-mode batch-norm), then ONE pair,
       which runs with the three batch-norm layers in eval mode (bce_optimizer.py:140-156).  Kept this short on purpose:
       the biases in front of a train-mode batch-norm random-walk on rounding noise (see tests/test_fit_conve.py), an
       eval-mode step reads them against lagging running means, and after a dozen steps that is enough to flip a relu
       gate in one implementation and not in another.
Stores the initial state dict once and the trained state dict (parameters + batch-norm running statistics) per run.

    python tests/golden/make_golden_fit_conve.py
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
from src.link_prediction.models import ConvE  # noqa: E402
from src.link_prediction.models.conve import ConvEHyperParams  # noqa: E402
from src.link_prediction.optimization import BCEOptimizer  # noqa: E402
from src.link_prediction.optimization.bce_optimizer import BCEOptimizerHyperParams  # noqa: E402

from tests.golden.make_golden import seed_all, synthetic_kg  # noqa: E402

MODEL_HP = dict(dimension=60, input_dropout_rate=0.0, feature_map_dropout_rate=0.0, hidden_dropout_rate=0.0, hidden_layer_size=1216)
HP = dict(batch_size=128, label_smoothing=0.1, lr=0.003, decay=0.995, epochs=3)


def n_pairs(ds):
    rows = np.vstack((ds.training_triples, ds.invert_triples(ds.training_triples)))
    return len({(int(s), int(p)) for s, p, _ in rows})


if __name__ == "__main__":
    n_ent, n_rel = 300, 10
    train, valid, test = synthetic_kg(7, n_ent, n_rel, 2400, 150, 150)
    refshim.register_dataset("golden-fit", train, valid, test, n_ent, n_rel)
    ds = Dataset("golden-fit")
    P = n_pairs(ds)
    bs_b = P - 1
    out = dict(train=train, valid=valid, test=test, n_ent=np.int64(n_ent), n_rel=np.int64(n_rel), n_pairs=np.int64(P),
               batch_b=np.int64(bs_b))
    seed_all(11)
    init = ConvE(ds, ConvEHyperParams(**MODEL_HP), init_random=True)
    with torch.no_grad():  # non-trivial affine batch-norm parameters so that their gradients matter from step one
        for bn in (init.batch_norm_1, init.batch_norm_2, init.batch_norm_3):
            bn.weight.add_(0.2 * torch.randn_like(bn.weight))
            bn.bias.add_(0.1 * torch.randn_like(bn.bias))
    state0 = {k: v.detach().clone() for k, v in init.state_dict().items()}
    for k, v in state0.items():
        out["init/" + k] = v.numpy()
    for tag, hp in (("a", HP), ("b", dict(HP, batch_size=bs_b, epochs=1))):
        model = ConvE(ds, ConvEHyperParams(**MODEL_HP), init_random=False)
        model.load_state_dict(state0, strict=False)
        if not hasattr(model, "entity_embeddings"):  # init_random=False leaves the tables unset in the reference
            model.entity_embeddings = torch.nn.Parameter(state0["entity_embeddings"].clone())
            model.relation_embeddings = torch.nn.Parameter(state0["relation_embeddings"].clone())
        seed_all(70)
        opt = BCEOptimizer(model=model, hp=BCEOptimizerHyperParams(**hp), verbose=False)
        opt.train(training_triples=ds.training_triples)
        for k, v in model.state_dict().items():
            if "num_batches_tracked" not in k:
                out[f"{tag}/" + k] = v.detach().numpy().copy()
        moved = {k: float((model.state_dict()[k] - state0[k]).abs().max()) for k in state0 if "num_batches" not in k}
        print(tag, "pairs", P, "batch", hp["batch_size"], "steps/epoch", -(-P // hp["batch_size"]), "last", P % hp["batch_size"])
        print("   moved:", {k: round(v, 5) for k, v in moved.items()})
    path = os.path.join(HERE, "conve_fit_small.npz")
    np.savez_compressed(path, **out)
    print("->", path, os.path.getsize(path), "bytes")
