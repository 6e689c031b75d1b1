"""Compact error report of the fused pass (kp_debug_contract) against fp64, per kernel variant."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kelpie_b200 import runtime
from tests.test_gpu_contract import _reference

def report(tag, ctx, q, ent, mode):
    m, l, O = ctx.contract(q, mode); torch.cuda.synchronize()
    m, l, O = [x.cpu().numpy().astype(np.float64) for x in (m, l, O)]
    rm, rl, rO, cond, _ = _reference(q, ent, mode)
    if mode == 0:
        sc = np.exp(rm - m); l, O = l / sc, O / sc[:, None]
    el = np.abs(l - rl) / np.abs(rl).max()
    eo = np.abs(O - rO).max(axis=1) / cond.max(axis=1)
    bad = np.argsort(-eo)[:4]
    print(f"{tag}: m-rm in [{(m-rm).min():.3g},{(m-rm).max():.3g}] l_err {el.max():.3g} O_err {eo.max():.3g} worst rows {bad.tolist()} {eo[bad]}", flush=True)

for (D, N, G) in [(512, 3001, 260), (512, 20011, 300), (256, 9001, 300)]:
    for mode in (0, 1):
        rng = np.random.default_rng(D + N + G + mode)
        ent = (rng.standard_normal((N, D)) * 0.3).astype(np.float32)
        q = (rng.standard_normal((G, D)) * (0.25 if mode == 0 else 0.05)).astype(np.float32)
        ctx = runtime.Context("ComplEx", ent, np.zeros((2, D), np.float32))
        for x4 in (1, 0):
            ctx.set_option("umma_x4", x4)
            report(f"D{D} N{N} G{G} mode{mode} x4={x4}", ctx, q, ent, mode)
        ctx.set_option("force_simt", 1)
        report(f"D{D} N{N} G{G} mode{mode} simt", ctx, q, ent, mode)
        ctx.close()
rng = np.random.default_rng(7)
N, D, G = 12001, 512, 256
ent = rng.standard_normal((N, D)).astype(np.float32) * 0.2
ent *= np.linspace(0.2, 3.0, N, dtype=np.float32)[:, None]
q = (rng.standard_normal((G, D)) * 0.6).astype(np.float32)
ctx = runtime.Context("ComplEx", ent, np.zeros((2, D), np.float32))
for x4 in (1, 0):
    ctx.set_option("umma_x4", x4)
    report(f"rescale x4={x4}", ctx, q, ent, 0)
ctx.set_option("force_simt", 1)
report("rescale simt", ctx, q, ent, 0)
