"""A few steps of the full-model ConvE trainer at DBpedia50 shape (for an ncu launch list)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kelpie_b200 import runtime
from kelpie_b200.data import Dataset
from kelpie_b200.link_prediction import MODEL_REGISTRY
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ds = Dataset.from_npz(os.path.join(ROOT, "tests", "golden", "dbpedia50_ids.npz"), name="DBpedia50")
cls, opt_cls = MODEL_REGISTRY["ConvE"]["class"], MODEL_REGISTRY["ConvE"]["optimizer"]
torch.manual_seed(0); np.random.seed(0)
m = cls(ds, cls.get_hyperparams_class()(dimension=200, input_dropout_rate=0.0, feature_map_dropout_rate=0.0, hidden_dropout_rate=0.0,
                                        hidden_layer_size=9728), init_random=True)
opt = opt_cls(model=m, hp=opt_cls.get_hyperparams_class()(batch_size=512, label_smoothing=0.1, lr=0.018, decay=0.995, epochs=1), verbose=False)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
orig = runtime.ConvEFit.steps
runtime.ConvEFit.steps = lambda self, order, off, lr, want_loss=False: orig(self, order, off[:n + 1], lr, want_loss)
opt.train(training_triples=ds.training_triples)
torch.cuda.synchronize()
print("ok", n, "steps")
