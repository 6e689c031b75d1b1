"""HBM roofline of the few-query all-entity pass (kp_stream.cu): TransE L2 scoring + fused
filtered rank of Q queries against a 1M x 256 fp32 table (1.02 GB, larger than the 126 MB L2).
Algorithmic bytes per launch = N*D*4 (the table is read once); prints achieved GB/s and the
fraction of MEASURED_PEAKS.json hbm_gbs.  Run on the GPU box: python tools/microbench_stream.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kelpie_b200 import runtime  # noqa: E402


def main():
    N, D, R2 = 1_000_000, 256, 64
    g = torch.Generator(device="cuda").manual_seed(1)
    ent = torch.randn(N, D, generator=g, device="cuda") * 0.05
    rel = torch.randn(R2, D, generator=g, device="cuda") * 0.05
    peak = 6650.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    out = []
    for kind, norm in (("TransE", 2), ("TransE", 1), ("ComplEx", 2)):
        ctx = runtime.Context(kind, ent, rel, norm=norm)
        for Q in (1, 2, 4, 8):
            rng = np.random.default_rng(Q)
            triples = np.stack([rng.integers(0, N, Q), rng.integers(0, R2, Q), rng.integers(0, N, Q)], 1)
            off = np.zeros(Q + 1, dtype=np.int64)
            tr = ctx.dev(triples, torch.int32)
            fo = ctx.dev(off, torch.int64)
            for _ in range(3):
                ctx.filtered_rank(tr, runtime.RANK_MODEL, flt_off=fo)
            ctx.set_option("timing", 1)
            ctx.stat("reset")
            for _ in range(20):
                ctx.filtered_rank(tr, runtime.RANK_MODEL, flt_off=fo)
            torch.cuda.synchronize()
            ms = ctx.stat("ms_pass") / ctx.stat("n_pass")
            ctx.set_option("timing", 0)
            gbs = N * D * 4 / (ms * 1e-3) / 1e9
            out.append(dict(model=kind, norm=norm, queries=Q, ms=round(ms, 4), achieved_gbs=round(gbs, 1),
                            frac_of_measured_hbm_peak=round(gbs / peak, 3)))
            print(out[-1], flush=True)
        ctx.close()
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "stream_microbench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
