#!/usr/bin/env python
"""bench.py -- candidate explanations evaluated per second (mimic post-training + filtered rank).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A *step* is one pass of the hot path over one batch of C synthetic candidate explanations
(plus the shared homologous "base" mimic): C+1 mimic post-trainings (all epochs) and the
filtered rank of the target for each.  Default workload = BASELINE.json configs[4]:
synthetic KG, 1M entities x ComplEx dim 256 (row = 512 fp32), 4096 candidates per batch per GPU
(weak scaling: every GPU holds the full tables and takes its own batch; one NCCL all-gather of
the (score, rank) pairs per step).  Other presets reproduce the shapes of configs[0..3].

`value` times only the kernels (inputs resident in HBM: CUDA events after the H2D copies and
before the D2H read); `e2e` times the same step from host numpy buffers through the C ABI
(pinned staging + H2D inside, D2H of scores/ranks inside).  `--impl reference` times the CPU
oracle port of the reference's own algorithm (torch CPU, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PRESETS = {
    # BASELINE.json configs[4] / SURVEY.md section 8(d) config 5
    "synthetic_complex_1m": dict(kind="ComplEx", N=1_000_000, dim=256, R=512, C=4096, T=(8, 64), init="normal0.1",
                                 hp=dict(optimizer_name="Adagrad", batch_size=512, epochs=43, lr=0.043, decay1=0.9,
                                         decay2=0.999, regularizer_name="N3", regularizer_weight=0)),
    # configs[1] shape (configs/ComplEx_DBpedia50_explanation.json)
    "complex_dbpedia50": dict(kind="ComplEx", N=24_620, dim=200, R=351, C=4096, T=(1, 12), init="normal0.1",
                              hp=dict(optimizer_name="Adagrad", batch_size=512, epochs=43, lr=0.043, decay1=0.9,
                                      decay2=0.999, regularizer_name="N3", regularizer_weight=0)),
    # configs[1] in sufficient mode: every candidate = 10 conversions (post_training_engine.py:178-191),
    # each its own mimic post-training (facts of the converted entity + the added rule) and rank
    "complex_dbpedia50_sufficient": dict(kind="ComplEx", N=24_620, dim=200, R=351, C=512, T=(2, 40), init="normal0.1", conversions=10,
                                         hp=dict(optimizer_name="Adagrad", batch_size=512, epochs=43, lr=0.043, decay1=0.9,
                                                 decay2=0.999, regularizer_name="N3", regularizer_weight=0)),
    # configs[0] shape (configs/TransE_DBpedia50_explanation.json)
    "transe_dbpedia50": dict(kind="TransE", N=24_620, dim=256, R=351, C=4096, T=(1, 11), init="xavier",
                             hp=dict(batch_size=2048, epochs=65, lr=0.01, margin=5, negative_triples_ratio=5,
                                     regularizer_weight=1.0)),
    # configs[3] shape (configs/TransE_YAGO4-20_explanation.json), synthetic stand-in for the missing train.txt
    "transe_yago4_20": dict(kind="TransE", N=96_000, dim=128, R=74, C=4096, T=(20, 60), init="xavier",
                            hp=dict(batch_size=2048, epochs=59, lr=0.01, margin=10, negative_triples_ratio=5,
                                    regularizer_weight=0.0)),
    # configs[2] shape (configs/ConvE_DB100K_explanation.json), hidden dropout 0.2 as in the config
    "conve_db100k": dict(kind="ConvE", N=99_604, dim=200, R=470, C=1024, T=(2, 20), init="xavier", dropout=(0.0, 0.0, 0.2),
                         hp=dict(batch_size=512, label_smoothing=0.1, lr=0.0432, decay=0.995, epochs=109)),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------- workload


def make_tables(cfg, device, seed=42):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    kind, N, R2 = cfg["kind"], cfg["N"], 2 * cfg["R"]
    D = 2 * cfg["dim"] if kind == "ComplEx" else cfg["dim"]
    if cfg["init"] == "normal0.1":
        ent = torch.randn(N, D, generator=g, device=device) * 0.1
        rel = torch.randn(R2, D, generator=g, device=device) * 0.1
    else:  # xavier_normal_ of the reference's random init (transe.py:27-33, conve.py:54-60)
        ent = torch.randn(N, D, generator=g, device=device) * (2.0 / (N + D)) ** 0.5
        rel = torch.randn(R2, D, generator=g, device=device) * (2.0 / (R2 + D)) ** 0.5
    conve = None
    if kind == "ConvE":
        H = D // 20
        hidden = 32 * 38 * (H - 2)
        cg = torch.Generator().manual_seed(seed)
        conve = dict(
            conv_w=torch.randn(32, 1, 3, 3, generator=cg) * 0.3, conv_b=torch.randn(32, generator=cg) * 0.1,
            fc_w=torch.randn(D, hidden, generator=cg) * (1.0 / hidden) ** 0.5, fc_b=torch.randn(D, generator=cg) * 0.1,
            dropout=tuple(cfg.get("dropout", (0.0, 0.0, 0.0))),
        )
        for i, n in ((1, 1), (2, 32), (3, D)):
            conve[f"bn{i}_w"] = torch.rand(n, generator=cg) * 0.5 + 0.75
            conve[f"bn{i}_b"] = torch.randn(n, generator=cg) * 0.1
            conve[f"bn{i}_mean"] = torch.randn(n, generator=cg) * 0.1
            conve[f"bn{i}_var"] = torch.rand(n, generator=cg) * 0.5 + 0.75
    return ent, rel, conve, D


def make_batch(cfg, D, C, seed):
    """C candidate jobs + 1 base job of one prediction family, as host numpy arrays."""
    from kelpie_b200 import plans

    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    kind, N, R = cfg["kind"], cfg["N"], cfg["R"]
    batch = plans.Batch(kind, N, R, cfg["hp"])
    jobs, filters = [], []
    tlo, thi = cfg["T"]
    for _ in range(C * int(cfg.get("conversions", 1)) + 1):
        T = int(rng.integers(tlo, thi + 1))
        x = rng.integers(0, N, size=T)
        r = rng.integers(0, R, size=T)
        head = rng.random(T) < 0.5
        facts = np.where(head[:, None], np.stack([np.full(T, N), r, x], 1), np.stack([x, r, np.full(T, N)], 1))
        if kind == "TransE":
            init = rng.standard_normal(D) * (2.0 / (D + 1)) ** 0.5
        elif kind == "ComplEx":
            init = rng.random(D) * 1e-3
        else:
            init = rng.random(D)
        batch.add(facts, init.astype(np.float32))
        jobs.append(facts)
        n_f = min(512, int(rng.zipf(2.0)))  # filter list: Zipf lengths (mean ~2, capped at 512)
        filters.append(np.unique(rng.integers(0, N, size=n_f)).astype(np.int32))
    arrs = batch.arrays(compact=not os.environ.get("KP_BENCH_FULL_TABLES"))  # TransE: compact index tables (kelpie_b200.h)
    p, o = int(rng.integers(0, R)), int(rng.integers(0, N))
    triples = np.tile(np.array([[N, p, o]], dtype=np.int32), (len(jobs), 1))
    flt_off = np.zeros(len(jobs) + 1, dtype=np.int64)
    flt_off[1:] = np.cumsum([len(f) for f in filters])
    flt_ids = np.concatenate(filters).astype(np.int32)
    return arrs, triples, flt_off, flt_ids, jobs, filters


def algorithmic_work(cfg, D, arrs):
    """executed algorithmic flops / bytes of ONE launch of the dominant kernel."""
    kind, N = cfg["kind"], cfg["N"]
    if kind == "TransE":
        # dominant kernel: the batched post-training (the L2 rank runs on tcgen05 since kp_rank_umma.cu took it over).
        # bytes = one corrupting row per training row + each candidate's own fact / relation rows once
        rows = float(arrs["row_off"][-1])
        facts = float(arrs["fact_off"][-1]) if "fact_off" in arrs else rows / cfg["hp"]["epochs"]
        return dict(bound="hbm", units=(rows + facts) * D * 4.0, what="transe_train")
    a_rows = int((arrs["pos"][:, 0] == N).sum())  # rows / pairs whose lhs is the mimic, per step
    return dict(bound="tensor", units=4.0 * a_rows * N * D, what="flash")


# ------------------------------------------------------------------------------- clocks


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), [x.strip() for x in line.split(",")]))

    def summary(self, t0, t1):
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for t, r in self.samples if t0 <= t <= t1 and len(r) >= 7] or [r for _, r in self.samples if len(r) >= 7]
        if not rows:
            return None
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows)}


# ------------------------------------------------------------------------------- reference arm


def run_reference(cfg, args, D, rank):
    """CPU oracle port of the reference's algorithm on a bounded sample: 1 candidate per step."""
    if rank != 0:
        return
    from oracle import kelpie_oracle as ko

    torch.set_num_threads(os.cpu_count() or 1)
    ent, rel, conve, _ = make_tables(cfg, "cpu")
    kind, N, R = cfg["kind"], cfg["N"], cfg["R"]
    kw = dict(norm=2, init_scale=1e-3)
    if kind == "ConvE":
        kw["conve"] = {k: v for k, v in conve.items() if k != "dropout"}
        kw["dropout"] = conve["dropout"]
    w = ko.Weights(kind, ent, rel, **kw)
    kg = ko.KG(np.zeros((0, 3), np.int64), np.zeros((0, 3), np.int64), np.zeros((0, 3), np.int64), N, R)
    n = args.warmup + args.steps
    arrs, triples, flt_off, flt_ids, jobs, filters = make_batch(cfg, D, max(n, 1), 1234)
    # Bounded sample: the whole run has to end within a few minutes whatever --steps / --warmup say.  Once the
    # wall-time budget is used up, warm-up candidates are cut short (a CPU has no clocks to ramp; the first
    # candidate pages the tables in) and the timed loop stops after the candidate in flight (>= 1 timed).
    budget = float(os.environ.get("KP_REFERENCE_BUDGET_S", 240.0))
    times, t_start, warm_done = [], time.perf_counter(), 0
    for i in range(n):
        spent = time.perf_counter() - t_start
        if i < args.warmup and i >= 1 and spent / i * (i + 1 + min(args.steps, 2)) > budget:
            continue  # not enough budget left for this warm-up candidate and (up to) two timed ones
        if i >= args.warmup and times and spent + spent / (warm_done + len(times)) > budget:
            break
        t0 = time.perf_counter()
        table = ko.post_train(w, kg, torch.from_numpy(arrs["init_rows"][i]).view(1, -1), jobs[i], cfg["hp"])
        ko.triple_results(w, table, tuple(int(x) for x in triples[i]), filters[i])
        dt = time.perf_counter() - t0
        log(f"[reference] candidate {i}: {dt:.2f} s")
        if i >= args.warmup:
            times.append(dt)
        else:
            warm_done += 1
    per = sum(times) / len(times) * int(cfg.get("conversions", 1))
    sample = (f"1 candidate per step ({len(times)} of {args.steps} steps timed after {warm_done} warm-up candidates, "
              f"wall-time budget {budget:.0f} s), T~U{cfg['T']} facts, all {cfg['hp']['epochs']} epochs + filtered rank")
    print(json.dumps({
        "impl": "reference", "metric": "candidate explanations evaluated/sec (post-train + filtered rank)",
        "value": 1.0 / per, "unit": "candidates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "steps_timed": len(times), "warmup_done": warm_done, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "model": kind, "entities": N, "row_floats": D, "relations": R,
                   "facts_per_candidate": list(cfg["T"]), "epochs": cfg["hp"]["epochs"],
                   "candidates_per_step": 1, "sample": "each step is ONE candidate of the workload's batch (bounded sample)"},
        "cpu_baseline": {"value": 1.0 / per, "unit": "candidates/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": 1.0 / per, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------- main


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="synthetic_complex_1m", choices=sorted(PRESETS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--candidates", type=int, default=None, help="override the preset's candidates per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="kp_set_option knob for A/B runs (e.g. umma_x4=0)")
    args = ap.parse_args()
    cfg = dict(PRESETS[args.workload])
    if args.candidates:
        cfg["C"] = args.candidates
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    D = 2 * cfg["dim"] if cfg["kind"] == "ComplEx" else cfg["dim"]

    if args.impl == "reference":
        run_reference(cfg, args, D, rank)
        return

    import torch.distributed as dist
    from kelpie_b200 import runtime

    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    kind, N, C = cfg["kind"], cfg["N"], cfg["C"]

    ent, rel, conve, D = make_tables(cfg, device)
    ctx = runtime.Context(kind, ent, rel, norm=2, conve=conve, device=local)
    for kv in args.opt:
        name, value = kv.split("=")
        ctx.set_option(name, int(value))
    hp = runtime.make_hp(kind, cfg["hp"])
    arrs, triples, flt_off, flt_ids, jobs, filters = make_batch(cfg, D, C, 1000 + rank)
    mode = runtime.RANK_ENGINE_MIN if kind == "TransE" else runtime.RANK_ENGINE_MAX
    work = algorithmic_work(cfg, D, arrs)
    dtypes = dict(init_rows=torch.float32, row_off=torch.int64, rows_per_epoch=torch.int32, pos=torch.int32,
                  neg=torch.int32, pos_off=torch.int64, pos_ids=torch.int32, fact_off=torch.int64, facts=torch.int32,
                  pos_idx=torch.uint16, neg_code=torch.int32)
    host = {k: v for k, v in arrs.items() if k != "static_epochs" and v is not None}
    h2d = sum(v.nbytes for v in host.values()) + triples.nbytes + flt_off.nbytes + flt_ids.nbytes
    max_rows, total_rows = int(arrs["rows_per_epoch"].max()), int(arrs["row_off"][-1])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device) if ent.numel() * 4 < (200 << 20) else None
    n_jobs = len(jobs)
    gathered = torch.empty((world, 2, n_jobs), dtype=torch.float32, device=device) if world > 1 else None

    def step():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        dev = {k: ctx.dev(v, dtypes[k]) for k, v in host.items()}
        tr, fo, fi = ctx.dev(triples, torch.int32), ctx.dev(flt_off, torch.int64), ctx.dev(flt_ids, torch.int32)
        ev[1].record()
        rows = ctx.post_train(hp, static_epochs=arrs["static_epochs"], max_rows_per_epoch=max_rows,
                              total_rows=total_rows, **dev)
        ts, bs, rk = ctx.filtered_rank(tr, mode, mimic_rows=rows, flt_off=fo, flt_ids=fi)
        ev[2].record()
        pair = torch.stack([ts, rk.to(torch.float32)])
        if world > 1:
            dist.all_gather_into_tensor(gathered, pair.unsqueeze(0))
            out = gathered.cpu()
        else:
            out = pair.cpu()
        ev[3].record()
        torch.cuda.synchronize()
        return ev, out

    def relevance(out):  # post_training_engine.py:136-145 on the host (C floats)
        sc, rk = out[..., 0, :].double().numpy(), out[..., 1, :].double().numpy()
        d = sc[..., 1:] - sc[..., :1] if kind == "TransE" else sc[..., :1] - sc[..., 1:]
        return (rk[..., 1:] - rk[..., :1]) + 1.0 / (1.0 + np.exp(-d))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # nvidia-smi starts before the warm-up: its start-up (fork, NVML initialisation) must not fall into a short timed region
    sampler = ClockSampler(local) if rank == 0 else None
    for i in range(args.warmup):
        t = time.time()
        step()
        if flush is not None:
            flush.fill_(i)
        log(f"[rank {rank}] warm-up step {i}: {time.time() - t:.2f} s")

    ctx.set_option("timing", 1)
    ctx.stat("reset")
    barrier()
    launches0, w0 = ctx.launches, time.time()
    t_kernel = t_e2e = 0.0
    for i in range(args.steps):
        ev, out = step()
        t_kernel += ev[1].elapsed_time(ev[2])
        t_e2e += ev[0].elapsed_time(ev[3])
        if flush is not None:
            flush.fill_(i)  # L2 flush between timed steps (outside the event brackets)
    barrier()
    w1 = time.time()
    launches = ctx.launches - launches0
    rel_vals = relevance(out)
    dom_ms, dom_n = ctx.stat("ms_" + work["what"]), ctx.stat("n_" + work["what"])
    breakdown = {c: round(ctx.stat("ms_" + c) / max(args.steps, 1), 3) for c in ("pass", "flash", "transe_train", "update", "conv")}
    ctx.set_option("timing", 0)

    t = torch.tensor([t_kernel, t_e2e], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_kernel, t_e2e = t.tolist()

    if rank == 0:
        total = C * world * args.steps
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        if work["bound"] == "tensor":
            peak, unit, src = peaks.get("bf16_tflops_sustained", 1400.0), "TFLOP/s", "bf16_tflops_sustained"
            achieved = work["units"] / (dom_ms / dom_n * 1e-3) / 1e12 if dom_n else None
        else:
            peak, unit, src = peaks.get("hbm_gbs", 6650.0), "GB/s", "hbm_gbs"
            achieved = work["units"] / (dom_ms / dom_n * 1e-3) / 1e9 if dom_n else None
        src += " of measured (MEASURED_PEAKS.json)" if peaks else " of fallback"
        traffic = None  # DRAM bytes per launch of the dominant kernel: from the committed ncu capture of this exact size
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))[f"{args.workload}/{C}"]["dram_bytes_per_launch"]
        except Exception:
            pass
        line = {
            "metric": "candidate explanations evaluated/sec (post-train + filtered rank)",
            "value": total / (t_kernel * 1e-3), "unit": "candidates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_kernel / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (bf16x3 split products, fp32 accumulate)" if work["bound"] == "tensor" else "f32",
            "data": "synthetic",
            "config": {"workload": args.workload, "model": kind, "entities": N, "row_floats": D, "relations": cfg["R"],
                       "candidates_per_step_per_gpu": C, "post_trainings_per_step_per_gpu": n_jobs,
                       "facts_per_candidate": list(cfg["T"]), "epochs": cfg["hp"]["epochs"],
                       "parallelism": f"candidate-sharded x{world}, tables replicated",
                       "l2": "tables exceed the 126 MB L2" if flush is None else "256 MB L2 flush between timed steps"},
            "e2e": {"value": total / (t_e2e * 1e-3), "unit": "candidates/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(out.numel() * 4), "ms_per_step": t_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": work["bound"], "kernel": work["what"], "achieved": achieved, "peak": peak, "unit": unit,
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": src,
                         "launches_timed": int(dom_n), "avg_launch_ms": (dom_ms / dom_n) if dom_n else None,
                         "share_of_step": dom_ms / t_kernel if t_kernel else None,
                         # fp32 parity on bf16 tensor cores: every product is hi*hi + hi*lo + lo*hi, so the
                         # tensor pipe executes 3 MMAs per algorithmic one and `frac` cannot exceed 1/3
                         **({"mma_per_product": 3, "executed_frac": 3 * achieved / peak if achieved else None}
                            if work["bound"] == "tensor" else {})},
            "kernel_ms_per_step": breakdown,  # CUDA-event time of the library's kernels by category (rank 0)
            "clocks": sampler.summary(w0, w1) if sampler else None,
            "relevance_checksum": float(np.nansum(rel_vals)),
        }
        if not args.no_cpu_baseline and world == 1:  # rank 0 at N = 1 only: the other ranks must not idle behind a CPU run
            line["cpu_baseline"] = cpu_baseline(cfg, D, ent, rel, conve)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(cfg, D, ent, rel, conve):
    """Oracle port timed on this box's host cores on a bounded sample (rank 0 only)."""
    from oracle import kelpie_oracle as ko

    torch.set_num_threads(os.cpu_count() or 1)
    kind, N, R = cfg["kind"], cfg["N"], cfg["R"]
    kw = dict(norm=2, init_scale=1e-3)
    if kind == "ConvE":
        kw["conve"] = {k: v.cpu() for k, v in conve.items() if k != "dropout"}
        kw["dropout"] = conve["dropout"]
    w = ko.Weights(kind, ent.cpu(), rel.cpu(), **kw)
    kg = ko.KG(np.zeros((0, 3), np.int64), np.zeros((0, 3), np.int64), np.zeros((0, 3), np.int64), N, R)
    budget, n, t_total = 20.0, 0, 0.0
    arrs, triples, flt_off, flt_ids, jobs, filters = make_batch(cfg, D, 16, 4321)
    while n < 16 and (n == 0 or t_total + t_total / n < budget):
        t0 = time.perf_counter()
        table = ko.post_train(w, kg, torch.from_numpy(arrs["init_rows"][n]).view(1, -1), jobs[n], cfg["hp"])
        ko.triple_results(w, table, tuple(int(x) for x in triples[n]), filters[n])
        t_total += time.perf_counter() - t0
        n += 1
    conv = int(cfg.get("conversions", 1))
    return {"value": n / t_total / conv, "unit": "candidates/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} mimic post-trainings (+ filtered rank) of the same workload, all epochs, {t_total:.1f} s of CPU time"
                      + (f"; one candidate = {conv} of them" if conv > 1 else "")}


if __name__ == "__main__":
    main()
