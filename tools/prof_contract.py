"""Time the fused pass alone (kp_debug_contract) on a config-5-shaped table and print where the MMA
thread of the cluster-4 kernel waits (option umma_prof)."""
import sys, os, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kelpie_b200 import runtime

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=1_000_000)
ap.add_argument("--D", type=int, default=512)
ap.add_argument("--G", type=int, default=16896)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
torch.manual_seed(0)
ent = (torch.randn(a.N, a.D, device="cuda") * 0.1)
q = (torch.randn(a.G, a.D, device="cuda") * 0.1)
ctx = runtime.Context("ComplEx", ent, torch.zeros(2, a.D, device="cuda"))
for kv in a.opt:
    n, v = kv.split("="); ctx.set_option(n, int(v))
ctx.contract(q, a.mode); torch.cuda.synchronize()
torch.cuda.synchronize()
ctx.set_option("umma_prof", 1)
ctx.set_option("timing", 1); ctx.stat("reset")
w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
w0.record()
for _ in range(a.reps):
    ctx.contract(q, a.mode)
w1.record()
torch.cuda.synchronize()
wall = w0.elapsed_time(w1) / a.reps  # whole call on the caller's stream (query split + every launch + merge)
ms = ctx.stat("ms_flash") / ctx.stat("n_flash")
fl = 4.0 * a.G * a.N * a.D
out = {"ms": ms, "tflops_alg": fl / ms / 1e9, "call_ms": wall, "call_tflops_alg": fl / wall / 1e9}
try:
    tot = ctx.stat("umma_prof_total")
    n_qt = (a.G + 255) // 256 * 2
    out["cycles_per_pair"] = tot / (n_qt * (a.reps + 0))  # pairs = 2 per cluster = n_qt (one per query tile)
    out["sm_mhz_in_kernel"] = out["cycles_per_pair"] / ms / 1e3
    for k in ("slot", "own", "for", "send", "wpin", "whdr", "wsfull", "soft"):
        out["wait_" + k] = ctx.stat("umma_prof_" + k) / max(tot, 1)
except RuntimeError:
    pass
try:  # S/V kernel (umma_x4 = 2): MMA-thread cycles of the scoring (S) and contracting (V) pairs
    n_cl = (a.G + 255) // 256
    for i, k in enumerate(("s_slot", "s_dep", "s_total", "v_slot", "v_dep", "v_total")):
        out[k + "_cycles_per_cluster"] = ctx.stat(f"umma_prof_{i}") / (n_cl * a.reps)
except RuntimeError:
    pass
print(out)
