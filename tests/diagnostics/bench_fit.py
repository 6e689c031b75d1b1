"""Full-model TransE training (kp_transe_fit_*) at configs/TransE_DBpedia50_training shape:
device time per epoch and per step, algorithmic HBM bytes of the dense Adam update, next to the oracle
restatement (= the reference's torch loop) on the host cores for a bounded sample of steps."""
import sys, os, json, time, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from kelpie_b200 import plans, runtime
from kelpie_b200.data import Dataset

ap = argparse.ArgumentParser()
ap.add_argument("--epochs", type=int, default=10)
ap.add_argument("--cpu-steps", type=int, default=20)
a = ap.parse_args()
ds = Dataset.from_npz(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dbpedia50_ids.npz"), name="DBpedia50")
N, R2, D = ds.num_entities, 2 * ds.num_relations, 256
hp = dict(batch_size=2048, lr=0.01, margin=5, negative_triples_ratio=5, regularizer_weight=1.0)
torch.manual_seed(0); np.random.seed(0)
ent = torch.nn.init.xavier_normal_(torch.empty(N, D)).cuda()
rel = torch.nn.init.xavier_normal_(torch.empty(R2, D)).cuda()
rows = np.vstack((ds.training_triples, ds.invert_triples(ds.training_triples))).astype(np.int64)
n = len(rows)
off = np.append(np.arange(0, n, hp["batch_size"]), n).astype(np.int64)
fit = runtime.TransEFit(ent, rel, 2, hp["lr"], hp["margin"], hp["regularizer_weight"])
draws = [plans.draw_transe_full_epoch(rows, N, 5) for _ in range(a.epochs + 1)]
dev = [(torch.from_numpy(p).cuda(), torch.from_numpy(q).cuda()) for p, q in draws]
fit.steps(dev[0][0], dev[0][1], off); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for p, q in dev[1:]:
    fit.steps(p, q, off)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
steps = a.epochs * (len(off) - 1)
adam_bytes = (N + R2) * D * 4 * 7  # g read + clear, m / v / p read + write
out = {"shape": {"entities": N, "relations2": R2, "dim": D, "rows_per_epoch": n, "steps_per_epoch": len(off) - 1},
       "gpu": {"ms_per_epoch": ms / a.epochs, "us_per_step": 1e3 * ms / steps, "steps_per_s": steps / (ms * 1e-3),
               "adam_algorithmic_GBps_if_all_time_were_adam": adam_bytes / (ms / steps * 1e-3) / 1e9}}
fit.close()
# CPU: the oracle restatement (torch autograd + torch.optim.Adam, dense [N, D] gradient) on the host cores
from oracle import kelpie_oracle as ko
torch.set_num_threads(os.cpu_count())
k = a.cpu_steps
t0 = time.perf_counter()
ko.train_transe_full(ent.cpu().numpy(), rel.cpu().numpy(), 2, ds.training_triples[: k * 1024], N, ds.num_relations,
                     dict(hp, epochs=1), n_epochs=1)
dt = time.perf_counter() - t0
out["cpu_port"] = {"steps": k, "us_per_step": 1e6 * dt / k, "steps_per_s": k / dt, "cores": os.cpu_count()}
out["speedup_steps_per_s"] = out["gpu"]["steps_per_s"] / out["cpu_port"]["steps_per_s"]
print(json.dumps(out))

# ---- ComplEx (configs/ComplEx_DBpedia50 shape: row 400, batch 512 against all 24 620 entities, Adagrad)
D = 400
chp = dict(optimizer_name="Adagrad", batch_size=512, lr=0.043, decay1=0.9, decay2=0.999, regularizer_name="N3", regularizer_weight=0)
torch.manual_seed(0)
ent = (torch.randn(N, D) * 0.1).cuda()
rel = (torch.randn(R2, D) * 0.1).cuda()
off = np.append(np.arange(0, n, 512), n).astype(np.int64)
cfit = runtime.ComplExFit(ent, rel, "Adagrad", chp["lr"], 0.9, 0.999, 0, 512)
perms = [torch.from_numpy(rows[torch.randperm(n).numpy()].astype(np.int32)).cuda() for _ in range(3)]
cfit.steps(perms[0], off); torch.cuda.synchronize()
e0.record()
for p in perms[1:]:
    cfit.steps(p, off)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
steps = 2 * (len(off) - 1)
flops = 3 * 2.0 * 512 * N * D  # logits, dQ = P E, gE = P^T Q
cout = {"shape": {"entities": N, "row": D, "batch": 512, "steps_per_epoch": len(off) - 1},
        "gpu": {"ms_per_epoch": ms / 2, "us_per_step": 1e3 * ms / steps, "steps_per_s": steps / (ms * 1e-3),
                "gemm_tflops_alg_if_all_time_were_gemm": flops / (ms / steps * 1e-3) / 1e12}}
cfit.close()
k = 6
t0 = time.perf_counter()
ko.train_complex_full(ent.cpu().numpy(), rel.cpu().numpy(), ds.training_triples[: k * 256], ds.num_relations, dict(chp, epochs=1), n_epochs=1)
dt = time.perf_counter() - t0
cout["cpu_port"] = {"steps": k, "us_per_step": 1e6 * dt / k, "steps_per_s": k / dt, "cores": os.cpu_count()}
cout["speedup_steps_per_s"] = cout["gpu"]["steps_per_s"] / cout["cpu_port"]["steps_per_s"]
print(json.dumps({"complex": cout}))

# ---- ConvE (configs/ConvE_DBpedia50 shape: dimension 200, hidden 9728, batch 512 pairs against all 24 620 entities, Adam)
from kelpie_b200.link_prediction import MODEL_REGISTRY
cls, opt_cls = MODEL_REGISTRY["ConvE"]["class"], MODEL_REGISTRY["ConvE"]["optimizer"]
torch.manual_seed(0); np.random.seed(0)
m = cls(ds, cls.get_hyperparams_class()(dimension=200, input_dropout_rate=0.0, feature_map_dropout_rate=0.0, hidden_dropout_rate=0.0,
                                        hidden_layer_size=9728), init_random=True)
vhp = dict(batch_size=512, label_smoothing=0.1, lr=0.018, decay=0.995, epochs=3)
opt = opt_cls(model=m, hp=opt_cls.get_hyperparams_class()(**vhp), verbose=False)
pairs, _, _ = opt_cls.er_vocab_tables(rows)
spe = -(-len(pairs) // 512)
# time the epochs' step calls only (CUDA events around every ConvEFit.steps call + host clock), not the per-train() setup
_orig_steps = runtime.ConvEFit.steps
_acc = {"ms": 0.0, "host_s": 0.0, "calls": 0, "pending": []}
def _timed_steps(self, order, off, lr, want_loss=False):
    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a_.record()
    out = _orig_steps(self, order, off, lr, want_loss)
    b_.record()
    _acc["host_s"] += time.perf_counter() - t0
    _acc["pending"].append((a_, b_))
    _acc["calls"] += 1
    return out
runtime.ConvEFit.steps = _timed_steps
opt.hp["epochs"] = 3
t0 = time.perf_counter()
opt.train(training_triples=ds.training_triples); torch.cuda.synchronize()
wall = time.perf_counter() - t0
runtime.ConvEFit.steps = _orig_steps
ev_ms = [x.elapsed_time(y) for x, y in _acc["pending"]]
ms = sum(ev_ms[1:])  # the first epoch grows the workspaces
steps = 2 * spe
flops = 2.0 * 512 * (3 * N * 200 + 3 * 200 * 9728)  # Z, dX, gE against the table; H, gW, dfeat through the Linear layer
vout = {"shape": {"entities": N, "dim": 200, "hidden": 9728, "batch": 512, "pairs": int(len(pairs)), "steps_per_epoch": spe},
        "gpu": {"ms_per_epoch": ms / 2, "us_per_step": 1e3 * ms / steps, "steps_per_s": steps / (ms * 1e-3),
                "gemm_tflops_alg_if_all_time_were_gemm": flops / (ms / steps * 1e-3) / 1e12,
                "host_issue_us_per_step": 1e6 * _acc["host_s"] / (3 * spe), "train_call_wall_s_3_epochs": wall,
                "timed": "CUDA events around the steps calls of epochs 2-3 (device + launch gaps), setup of train() excluded"}}
state = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items() if "num_batches" not in k}
k = 4
t0 = time.perf_counter()
ko.train_conve_full(state, ds.training_triples, N, ds.num_relations, vhp, n_epochs=1, max_steps=k)
dt = time.perf_counter() - t0
vout["cpu_port"] = {"steps": k, "us_per_step": 1e6 * dt / k, "steps_per_s": k / dt, "cores": os.cpu_count(), "includes": "er_vocab construction"}
vout["speedup_steps_per_s"] = vout["gpu"]["steps_per_s"] / vout["cpu_port"]["steps_per_s"]
print(json.dumps({"conve": vout}))
