"""Multi-GPU product path on real NCCL: one process per GPU, every rank holds the full tables,
`parallel.ShardedEngine` post-trains a cost-balanced slice of the candidates on its GPU and all-gathers the relevances.
Asserted: the gathered list equals the single-GPU list of the same engine (TransE: bit for bit; ComplEx / ConvE: to fp32
rounding, see the worker), every rank's generators end bit for bit where the single-process run leaves them, and the
StochasticBuilder driven through the sharded engine selects the same explanations with the same number of evaluated
candidates.  Needs >= 2 GPUs (`gpurun --gpus 2`); skipped elsewhere."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, random
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
kind = sys.argv[2]
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from tests.golden_util import load, seed_all
from tests.test_gpu_parity import _dataset, _model
from kelpie_b200.parallel import ShardedEngine
from kelpie_b200.relevance_engines import NecessaryPostTrainingEngine, SufficientPostTrainingEngine
from kelpie_b200.explanation_builders import StochasticBuilder

z, meta, kg, w, order = load(kind)
ds = _dataset(z)
for e, facts in order.items():
    ds.entity_to_training_triples[e] = [tuple(t) for t in facts]
m = _model(kind, z, meta, ds)
# one scoring kernel whatever the number of queries a rank ends up with (by default <= 8 queries take the streaming
# kernel, whose lane-strided sums differ from the tile kernel's sequential chain in the last bit)
m.context().set_option("force_tile", 1)
for case in meta["cases"]:
    cls = NecessaryPostTrainingEngine if case["mode"] == "necessary" else SufficientPostTrainingEngine
    pred = tuple(case["pred"])
    facts = [tuple(t) for t in order[pred[0]]]
    rules = [[f] for f in facts[:7]] + [[facts[0], facts[1]], [facts[1], facts[2], facts[3]]]
    out = {}
    for sharded in (False, True):
        eng = cls(m, ds, meta["hp"])
        seed_all(case["seed"])
        eng.set_cache()
        if case["mode"] == "sufficient":
            eng.select_entities_to_convert(pred, 3, 200)
        e = ShardedEngine(eng) if sharded else eng
        rels = e.compute_relevances(pred, rules)
        out[sharded] = (rels, float(torch.rand(1)), float(np.random.random()), float(torch.rand(1, device="cuda")))
    if out[True] != out[False]:
        print("last-bit differences", kind, case["tag"], "rank", dist.get_rank(), "\n sharded", out[True], "\n single ", out[False], flush=True)
    assert out[True][1:] == out[False][1:]   # the torch / numpy / CUDA generators: bit for bit
    # relevances: TransE bit for bit; ComplEx / ConvE to fp32 rounding -- the fused pass cuts the entity range into strips
    # according to how many rows the launch holds, so a slice sums the same terms in another order (integer rank deltas
    # are exact either way: a difference there would be >= 1 / base rank)
    if kind == "TransE":
        assert out[True][0] == out[False][0]
    else:
        np.testing.assert_allclose(out[True][0], out[False][0], rtol=2e-6, atol=1e-7)
# the explanation builder through the sharded engine: same explanation, same number of relevances
case = meta["cases"][0]
pred = tuple(case["pred"])
facts = [tuple(t) for t in order[pred[0]]][:6]
res = {}
for sharded in (False, True):
    eng = NecessaryPostTrainingEngine(m, ds, meta["hp"])
    seed_all(11); random.seed(11)
    eng.set_cache()
    b = StochasticBuilder(10.0 ** 9, ShardedEngine(eng) if sharded else eng, batch_size=4)
    r = b.build_explanations(pred, facts, k=5)
    res[sharded] = (r["rule_to_relevance"], r["#relevances"])
assert res[True][1] == res[False][1], (kind, res)                                # same number of evaluated candidates
assert [r for r, _ in res[True][0]] == [r for r, _ in res[False][0]], (kind, res)  # same explanations, same order
np.testing.assert_allclose([v for _, v in res[True][0]], [v for _, v in res[False][0]], rtol=2e-6, atol=1e-7)
dist.barrier()
dist.destroy_process_group()
print("ok")
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
@pytest.mark.parametrize("kind", ["TransE", "ComplEx", "ConvE"])
def test_sharded_engine_on_two_gpus_equals_one_gpu(kind, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT, kind], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=600)
        assert p.returncode == 0, out[-4000:]
