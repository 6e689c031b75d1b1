"""Phase timing of one TransE bench step (host wall clock and CUDA events per call): post_train and filtered_rank."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from kelpie_b200 import runtime

wl = sys.argv[1] if len(sys.argv) > 1 else "transe_dbpedia50"
variant = 0
cfg = bench.PRESETS[wl]

ent, rel, conve, D = bench.make_tables(cfg)
ctx = runtime.Context(cfg["kind"], ent.cuda(), rel.cuda(), norm=2, device=0)

hp = runtime.make_hp(cfg["kind"], cfg["hp"])
batch = bench.make_jobs(cfg, D, cfg["C"])
arrs, triples, flt_off, flt_ids = bench.build_arrays(cfg, batch, list(range(len(batch["jobs"]))))
dt = dict(init_rows=torch.float32, row_off=torch.int64, rows_per_epoch=torch.int32, pos=torch.int32, neg=torch.int32,
          fact_off=torch.int64, facts=torch.int32, pos_idx=torch.uint16, neg_code=torch.int32)
dev = {k: ctx.dev(v, dt[k]) for k, v in arrs.items() if k != "static_epochs" and v is not None}
tr, fo, fi = ctx.dev(triples, torch.int32), ctx.dev(flt_off, torch.int64), ctx.dev(flt_ids, torch.int32)
mr, tot = int(arrs["rows_per_epoch"].max()), int(arrs["row_off"][-1])
out = []
for it in range(6):
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t0 = time.perf_counter(); e[0].record()
    rows = ctx.post_train(hp, static_epochs=False, max_rows_per_epoch=mr, total_rows=tot, **dev)
    t1 = time.perf_counter(); e[1].record()
    ts, bs, rk = ctx.filtered_rank(tr, runtime.RANK_ENGINE_MIN, mimic_rows=rows, flt_off=fo, flt_ids=fi)
    t2 = time.perf_counter(); e[2].record()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    out.append(dict(host_post_train_ms=(t1 - t0) * 1e3, host_rank_ms=(t2 - t1) * 1e3, host_sync_ms=(t3 - t2) * 1e3,
                    dev_post_train_ms=e[0].elapsed_time(e[1]), dev_rank_ms=e[1].elapsed_time(e[2])))
print(json.dumps({"workload": wl, "variant": variant, "last": out[-1], "all_dev_post_train_ms": [round(o["dev_post_train_ms"], 3) for o in out]}))
