"""Filtered rank of Q ComplEx (DOT) or TransE (L2 distance, --kind TransE) queries against a config-5-shaped table: tensor-core pass with exact
re-check (kp_rank_umma.cu) vs the exact CUDA-core pass (kp_pass.cu).  Prints one JSON line."""
import sys, os, json, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kelpie_b200 import runtime

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=1_000_000)
ap.add_argument("--D", type=int, default=512)
ap.add_argument("--Q", type=int, default=4096)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--kind", default="ComplEx", choices=["ComplEx", "TransE"])
a = ap.parse_args()
torch.manual_seed(0)
ent = torch.randn(a.N, a.D, device="cuda") * 0.1
rel = torch.randn(16, a.D, device="cuda") * 0.1
ctx = runtime.Context(a.kind, ent, rel, norm=2)
rng = np.random.default_rng(0)
tr = np.stack([rng.integers(0, a.N, a.Q), rng.integers(0, 16, a.Q), rng.integers(0, a.N, a.Q)], 1).astype(np.int32)
off = np.arange(a.Q + 1, dtype=np.int64) * 2
ids = np.sort(rng.integers(0, a.N, (a.Q, 2)), axis=1).astype(np.int32).ravel()
out = {"kind": a.kind, "N": a.N, "D": a.D, "Q": a.Q, "flop_alg": (2.0 if a.kind == "ComplEx" else 3.0) * a.Q * a.N * a.D}
ranks = {}
for opt, name in ((1, "tcgen05"), (0, "cuda_core")):
    ctx.set_option("umma_rank", opt)
    ctx.filtered_rank(tr, 2, flt_off=off, flt_ids=ids); torch.cuda.synchronize()
    r0 = ctx.stat("rank_rechecks")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        ts, bs, rk = ctx.filtered_rank(tr, 2, flt_off=off, flt_ids=ids)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    ranks[opt] = rk.cpu().numpy()
    out[name] = {"ms": ms, "tflops_alg": out["flop_alg"] / ms / 1e9, "rechecks_per_call": (ctx.stat("rank_rechecks") - r0) / a.reps}
out["ranks_identical"] = bool(np.array_equal(ranks[0], ranks[1]))
print(json.dumps(out))
