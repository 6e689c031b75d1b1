"""Step-level comparison of the device ConvE trainer (kp_conve_fit_*) with the oracle restatement on the golden KG:
per-tensor max error relative to the tensor's max |.| after 1, 2, 5, 24 steps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tests.golden_util import GOLDEN, seed_all
from oracle import kelpie_oracle as ko
from kelpie_b200.data import Dataset
from kelpie_b200.link_prediction import MODEL_REGISTRY
from kelpie_b200 import runtime

z = np.load(os.path.join(GOLDEN, "conve_fit_small.npz"))
MODEL_HP = dict(dimension=60, input_dropout_rate=0.0, feature_map_dropout_rate=0.0, hidden_dropout_rate=0.0, hidden_layer_size=1216)
HP = dict(batch_size=int(sys.argv[1]) if len(sys.argv) > 1 else 128, label_smoothing=0.1, lr=0.003, decay=0.995, epochs=3)
ds = Dataset("golden-fit", z["train"], z["valid"], z["test"], int(z["n_ent"]), int(z["n_rel"]))
cls, opt_cls = MODEL_REGISTRY["ConvE"]["class"], MODEL_REGISTRY["ConvE"]["optimizer"]
state = {k: z["init/" + k] for k in ko.CONVE_STATE_KEYS}
for n_steps in [int(x) for x in (sys.argv[2].split(',') if len(sys.argv) > 2 else '1,2,5,24'.split(','))]:
    seed_all(70)
    want = ko.train_conve_full(state, z["train"], int(z["n_ent"]), int(z["n_rel"]), HP, max_steps=n_steps)
    m = cls(ds, cls.get_hyperparams_class()(**MODEL_HP), init_random=False)
    m.load_state_dict({k: torch.from_numpy(z["init/" + k]) for k in ko.CONVE_STATE_KEYS}, strict=False)
    seed_all(70)
    opt = opt_cls(model=m, hp=opt_cls.get_hyperparams_class()(**HP), verbose=True)
    orig = runtime.ConvEFit.steps
    budget = [n_steps]
    def limited(self, order, off, lr, want_loss=False):
        n = min(budget[0], len(off) - 1)
        budget[0] -= n
        return orig(self, order, off[:n + 1], lr, want_loss)
    runtime.ConvEFit.steps = limited
    opt.train(training_triples=ds.training_triples)
    runtime.ConvEFit.steps = orig
    got = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    print(f"--- {n_steps} steps, loss {opt.epoch_losses}")
    for k in ko.CONVE_STATE_KEYS:
        w = want[k]
        g = got[k].reshape(w.shape)
        moved = np.abs(w - state[k].reshape(w.shape)).max()
        if "-q" in sys.argv and k not in ("entity_embeddings", "relation_embeddings", "hidden_layer.weight", "convolutional_layer.weight", "batch_norm_3.running_mean"):
            continue
        print(f"  {k:32s} err {np.abs(g - w).max():.3e}  max {np.abs(w).max():.3e}  moved {moved:.3e}  rel-l2 {np.linalg.norm(g - w) / np.linalg.norm(w):.2e}")
