"""Isolate the operand layouts of the S/V kernel's contraction: q = 0 makes every probability 1 (A uniform: only the
B operand's layout matters); random q exercises both."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kelpie_b200 import runtime

rng = np.random.default_rng(0)
N, D, G = 2000, 512, 256
ent = (rng.standard_normal((N, D)) * 0.3).astype(np.float32)
ctx = runtime.Context("ComplEx", ent, np.zeros((2, D), np.float32))
ctx.set_option("umma_prof", 1)
e64 = ent.astype(np.float64)
for name, q in (("q=0", np.zeros((G, D), np.float32)), ("rand", (rng.standard_normal((G, D)) * 0.25).astype(np.float32))):
    z = q.astype(np.float64) @ e64.T
    m = z.max(1)
    P = np.exp(z - m[:, None])
    ref = P @ e64
    for dbg in [int(x) for x in sys.argv[1:]] or [0, 1, 2, 3]:
        ctx.set_option("sv_dbg", dbg)
        outs = []
        for rep in range(2):
            mm, ll, O = ctx.contract(q, 0)
            torch.cuda.synchronize()
            outs.append(O.cpu().numpy().astype(np.float64) * np.exp(mm.cpu().numpy().astype(np.float64) - m)[:, None])
        err = np.abs(outs[0] - ref)
        print(name, "dbg", dbg, "max err", err.max(), "ref scale", np.abs(ref).max(), "deterministic", np.array_equal(outs[0], outs[1]),
              "rows bad", int((err.max(1) > 1e-3 * np.abs(ref).max()).sum()), "cols bad", int((err.max(0) > 1e-3 * np.abs(ref).max()).sum()))
        rb = np.where(err.max(1) > 1e-3 * np.abs(ref).max())[0]
        print("  bad rows:", rb.tolist()[:80])
        if len(rb):
            r0 = rb[0]
            print("  row", r0, "O/ref first cols:", np.round(outs[0][r0, :8] / ref[r0, :8], 3).tolist(), "cols 64..71:", np.round(outs[0][r0, 64:72] / ref[r0, 64:72], 3).tolist(),
                  "cols 128..:", np.round(outs[0][r0, 128:132] / ref[r0, 128:132], 3).tolist(), "cols 256..:", np.round(outs[0][r0, 256:260] / ref[r0, 256:260], 3).tolist())
        if dbg & 16:
            print("   incomplete words at acquire: rows<96", ctx.stat("umma_prof_8"), "rows>=96", ctx.stat("umma_prof_9"), "rows checked", ctx.stat("umma_prof_10"))
        if dbg & 8:
            ratio = outs[0] / ref
            for r in (0, 1, 31, 32, 64, 95, 96, 97, 104, 112, 127, 128, 224):
                print("   row", r, "O/ref cols 0,1,2,300:", np.round(ratio[r, [0, 1, 2, 300]], 3).tolist())
        if dbg == 0:
            bad = np.argwhere(err > 1e-3 * np.abs(ref).max())
            print("  first bad (row, col):", bad[:6].tolist(), "bad cols sample:", sorted(set(bad[:, 1].tolist()))[:20])
