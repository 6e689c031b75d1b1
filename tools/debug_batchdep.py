"""Is a candidate's post-trained row independent of the batch it travels in?  Candidate 1 of the bench batch, post-trained
in a batch of 2 and in batches of 64 / 300 / 1200 candidates (1200: several waves of clusters), per kernel variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from kelpie_b200 import plans, runtime

cfg = dict(bench.PRESETS["synthetic_complex_1m"])
ent, rel, _, D = bench.make_tables(cfg)
N, R = cfg["N"], cfg["R"]
batch = bench.make_jobs(cfg, D, cfg["C"])
ctx = runtime.Context("ComplEx", ent.cuda(), rel.cuda())
E = int(sys.argv[1]) if len(sys.argv) > 1 else 43
hp = dict(cfg["hp"], epochs=E)
ref = {}
for name, opts in (("sv tps<=256", dict(umma_x4=2, umma_max_tps=256)), ("sv tps<=512", dict(umma_x4=2, umma_max_tps=512)), ("sv unbounded", dict(umma_x4=2, umma_max_tps=0))):
    for k, v in opts.items():
        ctx.set_option(k, v)
    for C in (2, 64, 300, 1200):
        b = plans.Batch("ComplEx", N, R, hp)
        for j in range(1, C + 1):
            b.add(batch["jobs"][j], batch["init_rows"][j])
        rows = ctx.post_train(runtime.make_hp("ComplEx", hp), **b.arrays()).cpu().numpy().astype(np.float64)
        if C == 2:
            ref[name] = rows[:2]
        d = np.abs(rows[:2] - ref[name]).max() / np.abs(ref[name]).max()
        print(f"{name:12s} C={C:5d}: first two candidates vs their C=2 rows: {d:.2e}", flush=True)

