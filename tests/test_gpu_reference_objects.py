"""the engines accept the REFERENCE'S OWN `Dataset` and model objects, and the unmodified
reference runs next to them on the same B200 (no CPU patch: `.cuda()` is real here), so this is a direct comparison
with the reference on the device: same objects, same seeds -> relevances within 1e-4 (integer rank deltas exact on
this 300-entity KG), for TransE (mimic row drawn by xavier_normal_ on the CUDA generator, as the reference on a GPU
does), ComplEx and ConvE.  Uses the copy of the reference staged under oracle/_ref (oracle/stage_ref.py); skipped
where that copy is absent.  Runs in a child process: the import shim registers stand-in modules (pykeen, optuna)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")

WORKER = r'''
import json, os, sys
root, kind = sys.argv[1], sys.argv[2]
sys.path.insert(0, root)
os.environ["KELPIE_REFERENCE_ROOT"] = os.path.join(root, "oracle", "_ref")
import numpy as np, torch
from oracle import refshim
refshim.install(cpu=False)                      # the unmodified reference, on the GPU
from tests.golden_util import GOLDEN, seed_all
from src.data import Dataset
from src.link_prediction.models import ComplEx, ConvE, TransE
from src.link_prediction.models.complex import ComplExHyperParams
from src.link_prediction.models.conve import ConvEHyperParams
from src.link_prediction.models.transe import TransEHyperParams
from src.relevance_engines import NecessaryPostTrainingEngine as RefNecessary
from kelpie_b200.relevance_engines import NecessaryPostTrainingEngine as FastNecessary

z = np.load(os.path.join(GOLDEN, f"{kind.lower()}_small.npz"))
meta = json.loads(bytes(z["meta"]).decode())
refshim.register_dataset("golden-objects", z["train"], z["valid"], z["test"], int(z["n_ent"]), int(z["n_rel"]))
ds = Dataset("golden-objects")
cls, hpc = {"TransE": (TransE, TransEHyperParams), "ComplEx": (ComplEx, ComplExHyperParams), "ConvE": (ConvE, ConvEHyperParams)}[kind]
m = cls(ds, hpc(**meta["params"]))
with torch.no_grad():
    m.entity_embeddings.copy_(torch.from_numpy(z["w_ent"]))
    m.relation_embeddings.copy_(torch.from_numpy(z["w_rel"]))
    if kind == "ConvE":
        m.convolutional_layer.weight.copy_(torch.from_numpy(z["w_conv_w"])); m.convolutional_layer.bias.copy_(torch.from_numpy(z["w_conv_b"]))
        m.hidden_layer.weight.copy_(torch.from_numpy(z["w_fc_w"])); m.hidden_layer.bias.copy_(torch.from_numpy(z["w_fc_b"]))
        for i, bn in enumerate((m.batch_norm_1, m.batch_norm_2, m.batch_norm_3), 1):
            bn.weight.copy_(torch.from_numpy(z[f"w_bn{i}_w"])); bn.bias.copy_(torch.from_numpy(z[f"w_bn{i}_b"]))
            bn.running_mean.copy_(torch.from_numpy(z[f"w_bn{i}_mean"])); bn.running_var.copy_(torch.from_numpy(z[f"w_bn{i}_var"]))
m.eval()
assert m.entity_embeddings.is_cuda
case = meta["cases"][0]
pred = tuple(case["pred"])
facts = [tuple(int(x) for x in t) for t in ds.entity_to_training_triples[pred[0]]]
rules = [[facts[0]], [facts[1]], [facts[0], facts[2]]]
out = {}
for name, engine_cls in (("reference", RefNecessary), ("fast", FastNecessary)):
    eng = engine_cls(m, ds, meta["hp"])
    seed_all(case["seed"])                      # CPU and CUDA generators
    eng.set_cache()
    out[name] = [float(eng.compute_relevance(pred, r)) for r in rules]
    torch.cuda.synchronize()
print("reference", out["reference"])
print("fast     ", out["fast"])
np.testing.assert_allclose(out["fast"], out["reference"], rtol=1e-4, atol=1e-4)
assert int(np.argmax(out["fast"])) == int(np.argmax(out["reference"]))
print("ok")
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference not staged (python oracle/stage_ref.py)")
@pytest.mark.parametrize("kind", ["TransE", "ComplEx", "ConvE"])
def test_engines_accept_the_reference_objects_and_match_the_reference_on_the_gpu(kind, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    p = subprocess.run([sys.executable, str(script), ROOT, kind], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, (p.stdout + p.stderr)[-4000:]
