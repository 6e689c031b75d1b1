"""43 epochs at 1M x 512 on the first candidates of the bench batch: which paths drift from the fp64 oracle, and when?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from oracle import kelpie_oracle as ko
from kelpie_b200 import plans, runtime

cfg = dict(bench.PRESETS["synthetic_complex_1m"])
ent, rel, _, D = bench.make_tables(cfg)
N, R = cfg["N"], cfg["R"]
batch = bench.make_jobs(cfg, D, cfg["C"])
ctx = runtime.Context("ComplEx", ent.cuda(), rel.cuda())
idx = [1, 2]
kg = ko.KG(np.zeros((0, 3), np.int64), np.zeros((0, 3), np.int64), np.zeros((0, 3), np.int64), N, R)
w32 = ko.Weights("ComplEx", ent, rel, init_scale=1e-3)
w64 = ko.Weights("ComplEx", ent.double(), rel.double(), init_scale=1e-3)
torch.set_num_threads(os.cpu_count())
for E in [int(x) for x in sys.argv[1:]] or [5, 10, 20, 43]:
    hp = dict(cfg["hp"], epochs=E)
    b = plans.Batch("ComplEx", N, R, hp)
    for j in idx:
        b.add(batch["jobs"][j], batch["init_rows"][j])
    arrs = b.arrays()
    rows = {}
    for name, opts in (("sv", dict(umma_x4=2, force_simt=0)), ("x4", dict(umma_x4=1, force_simt=0)), ("simt", dict(umma_x4=2, force_simt=1))):
        for k, v in opts.items():
            ctx.set_option(k, v)
        rows[name] = ctx.post_train(runtime.make_hp("ComplEx", hp), **arrs).cpu().numpy().astype(np.float64)
    init = lambda j, dt: torch.from_numpy(batch["init_rows"][j]).to(dt).view(1, -1)
    rows["o32"] = np.stack([ko.post_train(w32, kg, init(j, torch.float32), batch["jobs"][j], hp)[-1].numpy() for j in idx]).astype(np.float64)
    ref = np.stack([ko.post_train(w64, kg, init(j, torch.float64), batch["jobs"][j], hp)[-1].numpy() for j in idx])
    scale = np.abs(ref).max()
    d = {k: np.abs(v - ref).max() / scale for k, v in rows.items()}
    worst = np.unravel_index(np.argmax(np.abs(rows["sv"] - ref)), ref.shape)
    print(f"epochs {E}: vs fp64 oracle: " + "  ".join(f"{k} {v:.2e}" for k, v in d.items()), "| sv vs simt", f"{np.abs(rows['sv'] - rows['simt']).max() / scale:.2e}",
          "| worst component", worst, "sv", rows["sv"][worst], "simt", rows["simt"][worst], "o32", rows["o32"][worst], "o64", ref[worst], flush=True)
