"""Where does the 1M x 512 post-training deviate from the oracle?  Rows after 1 / 2 / 3 epochs: tcgen05 S/V kernel,
round-1 cluster kernel, CUDA-core pass, oracle (fp32 CPU) and an fp64 restatement of the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import kelpie_oracle as ko
from kelpie_b200 import plans, runtime

N, DIM, R = 1_000_000, 256, 512
g = torch.Generator(device="cuda").manual_seed(42)
ent = torch.randn(N, 2 * DIM, generator=g, device="cuda") * 0.1
rel = torch.randn(2 * R, 2 * DIM, generator=g, device="cuda") * 0.1
ctx = runtime.Context("ComplEx", ent, rel)
rng = np.random.default_rng(11)
torch.manual_seed(0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 60
jobs, inits = [], []
for _ in range(3):
    x, r = rng.choice(N, T, replace=False), rng.integers(0, R, T)
    jobs.append(np.stack([np.full(T, N), r, x], 1))
    inits.append((rng.random(2 * DIM) * 1e-3).astype(np.float32))
ent_h, rel_h = ent.cpu(), rel.cpu()
kg = ko.KG(np.zeros((0, 3), np.int64), np.zeros((0, 3), np.int64), np.zeros((0, 3), np.int64), N, R)
for E in (1, 2, 3):
    hp = dict(optimizer_name="Adagrad", batch_size=512, epochs=E, lr=0.043, decay1=0.9, decay2=0.999, regularizer_name="N3", regularizer_weight=0)
    b = plans.Batch("ComplEx", N, R, hp)
    for f, i in zip(jobs, inits):
        b.add(f, i)
    arrs = b.arrays()
    rows = {}
    for name, opts in (("sv", dict(umma_x4=2, force_simt=0)), ("x4", dict(umma_x4=1, force_simt=0)), ("pair", dict(umma_x4=0, force_simt=0)), ("simt", dict(umma_x4=2, force_simt=1))):
        for k, v in opts.items():
            ctx.set_option(k, v)
        rows[name] = ctx.post_train(runtime.make_hp("ComplEx", hp), **arrs).cpu().numpy().astype(np.float64)
    w = ko.Weights("ComplEx", ent_h, rel_h, init_scale=1e-3)
    rows["oracle"] = np.stack([ko.post_train(w, kg, torch.from_numpy(inits[c]).view(1, -1), jobs[c], hp)[-1].numpy() for c in range(3)]).astype(np.float64)
    if E == 1:
        w64 = ko.Weights("ComplEx", ent_h.double(), rel_h.double(), init_scale=1e-3)
        try:
            rows["oracle64"] = np.stack([ko.post_train(w64, kg, torch.from_numpy(inits[c]).double().view(1, -1), jobs[c], hp)[-1].numpy() for c in range(3)])
        except Exception as e:
            print("fp64 oracle failed:", e)
    ref = rows["oracle"]
    scale = np.abs(ref).max(axis=1, keepdims=True)
    print(f"epochs {E}: " + "  ".join(f"{k}: {np.abs(v - ref).max() / scale.max():.2e}" for k, v in rows.items() if k != "oracle"),
          "| sv vs simt", f"{np.abs(rows['sv'] - rows['simt']).max() / scale.max():.2e}", "| moved", f"{np.abs(ref - np.stack(inits)).max():.2e}")
    if E == 1:
        d = np.abs(rows["sv"] - ref) / scale
        bad = np.argwhere(d > 1e-4)
        print("  components off by > 1e-4:", len(bad), "first:", [(int(c), int(k), float(rows['sv'][c, k]), float(ref[c, k]), float(inits[c][k])) for c, k in bad[:6]])
