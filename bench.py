#!/usr/bin/env python
"""bench.py -- candidate explanations evaluated per second (mimic post-training + filtered rank).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A *step* is one pass of the hot path over ONE batch of C synthetic candidate explanations (plus the shared
homologous "base" mimic): C+1 mimic post-trainings (all epochs) and the filtered rank of the target for
each.  Default workload = BASELINE.json configs[4]: synthetic KG, 1M entities x ComplEx dim 256 (row =
512 fp32), one prediction <s, p, o> whose subject has 72 training facts, 4096 candidate explanations in
necessary mode (each removes a subset so that T ~ U{8..64} facts remain).  With --gpus N the SAME batch is
cut into N contiguous slices of near-equal cost (kelpie_b200.parallel.shard_bounds over the facts of each
candidate), every rank holds the full tables and post-trains its slice (+ the base mimic, redundantly, as
SURVEY.md 8e prescribes), and one NCCL all-gather collects the (score, rank) pairs: strong scaling.
The other presets reproduce the shapes of configs[0..3].

`value` times only the kernels (inputs resident in HBM: CUDA events after the H2D copies and before the D2H
read); `e2e` times the same step from host numpy buffers through the C ABI (pinned staging + H2D inside,
D2H of scores / ranks inside).  `cpu_baseline`: the oracle port of the reference's algorithm, run by a child
process on the host cores WHILE the GPU warms up, on the first candidates of the same batch; its rows / ranks
are compared with the GPU's (`parity`).  `--impl reference` times the reference's own CPU code: the
unmodified `NecessaryPostTrainingEngine.compute_relevance` staged under oracle/_ref (oracle/stage_ref.py)
when present and the workload is a prediction family, else the oracle port.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

T_PROCESS_START = time.time()

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PRESETS = {
    # BASELINE.json configs[4] / SURVEY.md section 8(d) config 5
    "synthetic_complex_1m": dict(kind="ComplEx", N=1_000_000, dim=256, R=512, C=4096, T=(8, 64), family=72, init="normal0.1",
                                 hp=dict(optimizer_name="Adagrad", batch_size=512, epochs=43, lr=0.043, decay1=0.9,
                                         decay2=0.999, regularizer_name="N3", regularizer_weight=0)),
    # the same prediction-family generator at a size the CPU tests finish in seconds (not a BASELINE config)
    "synthetic_complex_small": dict(kind="ComplEx", N=3000, dim=16, R=8, C=64, T=(4, 12), family=16, init="normal0.1",
                                    hp=dict(optimizer_name="Adagrad", batch_size=512, epochs=5, lr=0.043, decay1=0.9,
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
BATCH_SEED = 1000


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------- workload (numpy / torch CPU only)


def make_tables(cfg, seed=42):
    """Entity / relation tables (and the frozen ConvE network) on the HOST from a seeded CPU generator, so that
    every arm -- GPU ranks, the CPU leg, the reference arm -- holds bit-identical weights."""
    g = torch.Generator()
    g.manual_seed(seed)
    kind, N, R2 = cfg["kind"], cfg["N"], 2 * cfg["R"]
    D = 2 * cfg["dim"] if kind == "ComplEx" else cfg["dim"]
    if cfg["init"] == "normal0.1":
        ent = torch.randn(N, D, generator=g) * 0.1
        rel = torch.randn(R2, D, generator=g) * 0.1
    else:  # xavier_normal_ of the reference's random init (transe.py:27-33, conve.py:54-60)
        ent = torch.randn(N, D, generator=g) * (2.0 / (N + D)) ** 0.5
        rel = torch.randn(R2, D, generator=g) * (2.0 / (R2 + D)) ** 0.5
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


def _init_row(kind, D, rng):
    if kind == "TransE":
        return (rng.standard_normal(D) * (2.0 / (D + 1)) ** 0.5).astype(np.float32)
    if kind == "ComplEx":
        return (rng.random(D) * 1e-3).astype(np.float32)
    return rng.random(D).astype(np.float32)


def make_jobs(cfg, D, C, seed=BATCH_SEED):
    """The batch as host data: job 0 = the homologous (base) mimic, jobs 1..C = the candidates.

    Prediction family (preset key `family` = F): one prediction <s, p, o>; s has F training facts; candidate c
    removes a random subset so that T_c ~ U{tlo..thi} facts remain (necessary mode, post_training_engine.py:147-158);
    the filter of every job is the reference's to_filter[(mimic, p)]: o itself (a test fact) and the objects of the
    remaining facts (mimic, p, x) (kelpie_dataset.py:57-62,145-153).
    Otherwise: independent random fact sets per job and Zipf filter lists (the round-1 generator)."""
    rng = np.random.default_rng(seed)
    kind, N, R = cfg["kind"], cfg["N"], cfg["R"]
    tlo, thi = cfg["T"]
    n_jobs = C * int(cfg.get("conversions", 1)) + 1
    out = dict(jobs=[], filters=[], init_rows=np.empty((n_jobs, D), np.float32))
    if cfg.get("family"):
        F = int(cfg["family"])
        s = int(rng.integers(0, N))
        x = rng.choice(N - 1, size=F, replace=False)
        x[x >= s] += 1
        r = rng.integers(0, R, size=F)
        head = rng.random(F) < 0.5
        facts_s = np.where(head[:, None], np.stack([np.full(F, s), r, x], 1), np.stack([x, r, np.full(F, s)], 1)).astype(np.int64)
        p = int(rng.integers(0, R))
        o = int(rng.integers(0, N - 1))
        o += o >= s
        mimic = np.where(facts_s == s, N, facts_s)
        mimic[:, 1] = facts_s[:, 1]
        keeps = [np.arange(F)] + [np.sort(rng.choice(F, size=int(rng.integers(tlo, thi + 1)), replace=False)) for _ in range(n_jobs - 1)]
        init = _init_row(kind, D, rng)  # one torch.rand(1, D) per compute_relevance; the bench draws one per job all the same
        for j, keep in enumerate(keeps):
            fj = mimic[keep]
            out["jobs"].append(fj)
            objs = fj[(fj[:, 0] == N) & (fj[:, 1] == p), 2]
            out["filters"].append(np.unique(np.concatenate([objs, [o]])).astype(np.int32))
            out["init_rows"][j] = init if j == 0 else _init_row(kind, D, rng)
        out.update(pred=(s, p, o), facts_s=facts_s, keeps=keeps, triple=(N, p, o))
    else:
        for j in range(n_jobs):
            T = int(rng.integers(tlo, thi + 1))
            x = rng.integers(0, N, size=T)
            r = rng.integers(0, R, size=T)
            head = rng.random(T) < 0.5
            out["jobs"].append(np.where(head[:, None], np.stack([np.full(T, N), r, x], 1), np.stack([x, r, np.full(T, N)], 1)))
            out["init_rows"][j] = _init_row(kind, D, rng)
            n_f = min(512, int(rng.zipf(2.0)))  # filter list: Zipf lengths (mean ~2, capped at 512)
            out["filters"].append(np.unique(rng.integers(0, N, size=n_f)).astype(np.int32))
        out["triple"] = (N, int(rng.integers(0, R)), int(rng.integers(0, N)))
    return out


def build_arrays(cfg, batch_jobs, idx, seed=BATCH_SEED):
    """kp_pt_batch arrays of the jobs `idx` (imports kelpie_b200: GPU arm only)."""
    from kelpie_b200 import plans

    np.random.seed(seed)
    torch.manual_seed(seed)
    b = plans.Batch(cfg["kind"], cfg["N"], cfg["R"], cfg["hp"])
    for j in idx:
        b.add(batch_jobs["jobs"][j], batch_jobs["init_rows"][j])
    arrs = b.arrays(compact=not os.environ.get("KP_BENCH_FULL_TABLES"))  # TransE: compact index tables (kelpie_b200.h)
    filters = [batch_jobs["filters"][j] for j in idx]
    flt_off = np.zeros(len(idx) + 1, dtype=np.int64)
    flt_off[1:] = np.cumsum([len(f) for f in filters])
    flt_ids = np.concatenate(filters).astype(np.int32)
    triples = np.tile(np.array([batch_jobs["triple"]], dtype=np.int32), (len(idx), 1))
    return arrs, triples, flt_off, flt_ids


def algorithmic_work(cfg, D, arrs):
    """executed algorithmic flops / bytes of ONE launch of the dominant kernel."""
    kind, N = cfg["kind"], cfg["N"]
    if kind == "TransE":
        # dominant kernel: the batched post-training.  It is latency / issue bound (DRAM idle, issue
        # slots 66-77 % busy); the gather rate is reported for reference, not as an HBM roofline claim.
        rows = float(arrs["row_off"][-1])
        facts = float(arrs["fact_off"][-1]) if "fact_off" in arrs else rows / cfg["hp"]["epochs"]
        return dict(bound="latency", units=(rows + facts) * D * 4.0, what="transe_train")
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


# ------------------------------------------------------------------------------- CPU arms (no kelpie_b200 import)


def _oracle_setup(cfg):
    from oracle import kelpie_oracle as ko

    torch.set_num_threads(os.cpu_count() or 1)
    ent, rel, conve, D = make_tables(cfg)
    kind, N, R = cfg["kind"], cfg["N"], cfg["R"]
    kw = dict(norm=2, init_scale=1e-3)
    if kind == "ConvE":
        kw["conve"] = {k: v for k, v in conve.items() if k != "dropout"}
        kw["dropout"] = conve["dropout"]
    w = ko.Weights(kind, ent, rel, **kw)
    kg = ko.KG(np.zeros((0, 3), np.int64), np.zeros((0, 3), np.int64), np.zeros((0, 3), np.int64), N, R)
    return ko, w, kg, D, float(ent.double().sum())


def _oracle_candidate(ko, w, kg, cfg, batch, j):
    """Oracle port: post-train job j and rank its target; returns (seconds, row, score, rank)."""
    t0 = time.perf_counter()
    table = ko.post_train(w, kg, torch.from_numpy(batch["init_rows"][j]).view(1, -1), batch["jobs"][j], cfg["hp"])
    res = ko.triple_results(w, table, tuple(int(x) for x in batch["triple"]), batch["filters"][j])
    dt = time.perf_counter() - t0
    row = table[-1].detach().numpy().astype(np.float64) if isinstance(table, torch.Tensor) else None
    return dt, row, float(res["target_score"]), int(res["target_rank"])


def run_cpu_leg(cfg, args):
    """Child process of the GPU arm (rank 0, N = 1): the oracle port on the first candidates of the SAME batch, on the
    host cores, while the GPU warms up.  Results are rewritten to args.cpu_leg after every candidate."""
    ko, w, kg, D, checksum = _oracle_setup(cfg)
    batch = make_jobs(cfg, D, cfg["C"])
    budget = float(os.environ.get("KP_CPU_LEG_BUDGET_S", 90.0))
    out = dict(ent_checksum=checksum, cores=torch.get_num_threads(), times=[], rows=[], scores=[], ranks=[], jobs=[])

    def dump():
        tmp = args.cpu_leg + ".tmp"
        with open(tmp, "w") as f:
            json.dump(out, f)
        os.replace(tmp, args.cpu_leg)

    total = 0.0
    if cfg["kind"] == "TransE":
        # the negatives / shuffles come from the torch and numpy generators in job order (build_arrays seeds them and adds
        # job 0, 1, 2, ...): consume job 0's draws first so that candidate j sees the numbers the GPU arm's plan holds
        np.random.seed(BATCH_SEED)
        torch.manual_seed(BATCH_SEED)
        _oracle_candidate(ko, w, kg, cfg, batch, 0)
    for j in range(1, min(9, len(batch["jobs"]))):
        if out["times"] and len(out["times"]) >= 2 and total + total / len(out["times"]) > budget:
            break
        dt, row, score, rank = _oracle_candidate(ko, w, kg, cfg, batch, j)
        total += dt
        out["times"].append(dt)
        out["rows"].append(None if row is None else row.tolist())
        out["scores"].append(score)
        out["ranks"].append(rank)
        out["jobs"].append(j)
        dump()
    if cfg["kind"] != "ComplEx":
        return
    # the reference arithmetic's own reproducibility at this size: the first candidate again in fp64 (untimed; ComplEx
    # only -- its single-step epochs do not depend on the drawn permutations)
    try:
        w64 = ko.Weights(cfg["kind"], w.ent.double(), w.rel.double(), init_scale=1e-3)
        t64 = ko.post_train(w64, kg, torch.from_numpy(batch["init_rows"][1]).double().view(1, -1), batch["jobs"][1], cfg["hp"])
        out["row64"] = t64[-1].detach().numpy().tolist()
        dump()
    except Exception as e:  # ConvE weights etc.: the fp64 restatement is optional
        out["row64_error"] = repr(e)
        dump()


def _reference_line(cfg, args, D, per, times, warm_done, kind_tag, cores, sample):
    n = len(times)
    return {
        "impl": "reference", "metric": "candidate explanations evaluated/sec (post-train + filtered rank)",
        "value": 1.0 / per, "unit": "candidates/s", "n_gpus": args.gpus, "steps": n, "warmup": warm_done,
        "steps_requested": args.steps, "warmup_requested": args.warmup, "ms_per_step": per * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "model": cfg["kind"], "entities": cfg["N"], "row_floats": D, "relations": cfg["R"],
                   "facts_per_candidate": list(cfg["T"]), "epochs": cfg["hp"]["epochs"], "candidates_per_step": 1,
                   "sample": "each step is ONE candidate of the workload's batch (bounded sample of the same workload)"},
        "cpu_baseline": {"value": 1.0 / per, "unit": "candidates/s", "cores": cores, "kind": kind_tag, "sample": sample},
        "e2e": {"value": 1.0 / per, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def _bounded_loop(args, budget, one):
    """Warm-up and timed candidates under a wall-time budget: once it is used up, warm-up candidates are cut short
    (a CPU has no clocks to ramp; the first candidate pages the tables in) and the timed loop stops after the
    candidate in flight.  `steps` of the printed line is the number actually timed."""
    times, t_start, warm_done = [], time.perf_counter(), 0
    min_timed = min(args.steps, 4)
    for i in range(args.warmup + args.steps):
        spent = time.perf_counter() - t_start
        done = warm_done + len(times)
        if i < args.warmup and done >= 1 and spent / done * (done + 1 + min_timed) > budget:
            continue  # not enough budget left for this warm-up candidate and the timed ones
        if i >= args.warmup and len(times) >= 1 and spent + spent / done > budget:
            break
        dt = one(i)
        log(f"[reference] candidate {i}: {dt:.2f} s")
        if i >= args.warmup:
            times.append(dt)
        else:
            warm_done += 1
    return times, warm_done


def run_reference(cfg, args, D, rank):
    """The reference's own CPU path on a bounded sample: 1 candidate per step, all host threads."""
    if rank != 0:
        return
    budget = float(os.environ.get("KP_REFERENCE_BUDGET_S", 400.0))
    ref_src = os.path.join(ROOT, "oracle", "_ref", "src")
    if cfg.get("family") and os.path.isdir(ref_src) and not os.environ.get("KP_REFERENCE_PORT"):
        return run_reference_staged(cfg, args, D, budget)
    ko, w, kg, D, _ = _oracle_setup(cfg)
    batch = make_jobs(cfg, D, max(args.warmup + args.steps, 1))
    times, warm_done = _bounded_loop(args, budget, lambda i: _oracle_candidate(ko, w, kg, cfg, batch, 1 + i)[0])
    per = sum(times) / len(times) * int(cfg.get("conversions", 1))
    sample = (f"1 candidate per step ({len(times)} timed after {warm_done} warm-up candidates, wall-time budget {budget:.0f} s), "
              f"T~U{cfg['T']} facts, all {cfg['hp']['epochs']} epochs + filtered rank; oracle port (oracle/kelpie_oracle.py)")
    print(json.dumps(_reference_line(cfg, args, D, per, times, warm_done, "port", torch.get_num_threads(), sample)))


def run_reference_staged(cfg, args, D, budget):
    """The UNMODIFIED reference (oracle/_ref, staged by oracle/stage_ref.py) through its own public API:
    NecessaryPostTrainingEngine.compute_relevance(pred, removed facts) on the CPU (oracle/refshim.py redirects the
    hard-coded .cuda() calls and supplies the three absent third-party imports)."""
    os.environ["KELPIE_REFERENCE_ROOT"] = os.path.join(ROOT, "oracle", "_ref")
    from oracle import refshim

    refshim.install(cpu=True)
    torch.set_num_threads(os.cpu_count() or 1)
    ent, rel, _, D = make_tables(cfg)
    N, R = cfg["N"], cfg["R"]
    batch = make_jobs(cfg, D, max(args.warmup + args.steps, 1))
    s, p, o = batch["pred"]
    facts_s = batch["facts_s"]
    refshim.register_dataset("kp_bench_family", facts_s, np.zeros((0, 3), np.int64), np.array([[s, p, o]], np.int64), N, R)
    from src.data import Dataset
    from src.link_prediction.models import ComplEx
    from src.link_prediction.models.complex import ComplExHyperParams
    from src.relevance_engines import NecessaryPostTrainingEngine

    assert cfg["kind"] == "ComplEx", "the staged reference arm covers the ComplEx prediction-family workload"
    t0 = time.perf_counter()
    ds = Dataset("kp_bench_family")
    model = ComplEx(ds, ComplExHyperParams(dimension=cfg["dim"], init_scale=1e-3), init_random=True)
    with torch.no_grad():
        model.entity_embeddings.copy_(ent)
        model.relation_embeddings.copy_(rel)
    model.eval()
    eng = NecessaryPostTrainingEngine(model, ds, cfg["hp"])
    eng.set_cache()
    log(f"[reference] dataset + model + engine: {time.perf_counter() - t0:.1f} s")
    torch.manual_seed(BATCH_SEED)
    np.random.seed(BATCH_SEED)

    def one(i):
        keep = set(batch["keeps"][1 + i].tolist())
        rule = [tuple(int(v) for v in facts_s[f]) for f in range(len(facts_s)) if f not in keep]
        t = time.perf_counter()
        eng.compute_relevance((s, p, o), rule)
        return time.perf_counter() - t

    times, warm_done = _bounded_loop(args, budget, one)
    per = sum(times) / len(times)
    sample = (f"1 candidate per step ({len(times)} timed after {warm_done} warm-up candidates -- the first also post-trains the shared base "
              f"mimic and deep-copies the dataset -- wall-time budget {budget:.0f} s), T~U{cfg['T']} of {cfg['family']} facts kept, all "
              f"{cfg['hp']['epochs']} epochs + filtered rank; the unmodified reference's NecessaryPostTrainingEngine.compute_relevance "
              f"(oracle/_ref, CPU-patched)")
    print(json.dumps(_reference_line(cfg, args, D, per, times, warm_done, "reference", torch.get_num_threads(), sample)))


# ------------------------------------------------------------------------------- GPU arm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="synthetic_complex_1m", choices=sorted(PRESETS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--candidates", type=int, default=None, help="override the preset's candidates per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-leg", default=None, help=argparse.SUPPRESS)  # internal: child process of the GPU arm
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="kp_set_option knob for A/B runs (e.g. umma_x4=1)")
    args = ap.parse_args()
    cfg = dict(PRESETS[args.workload])
    if args.candidates:
        cfg["C"] = args.candidates
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    D = 2 * cfg["dim"] if cfg["kind"] == "ComplEx" else cfg["dim"]

    if args.cpu_leg:
        run_cpu_leg(cfg, args)
        return
    if args.impl == "reference":
        run_reference(cfg, args, D, rank)
        return

    # the CPU leg starts first: it overlaps table generation and the GPU warm-up (the host is otherwise idle)
    leg, leg_path = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        leg_path = os.path.join(tempfile.mkdtemp(prefix="kp_bench_"), "cpu_leg.json")
        cmd = [sys.executable, os.path.abspath(__file__), "--cpu-leg", leg_path, "--workload", args.workload]
        if args.candidates:
            cmd += ["--candidates", str(args.candidates)]
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
        leg = subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env)

    import torch.distributed as dist
    from kelpie_b200 import parallel, runtime

    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    kind, N, C = cfg["kind"], cfg["N"], cfg["C"]
    conv = int(cfg.get("conversions", 1))

    ent_h, rel_h, conve, D = make_tables(cfg)
    ent_checksum = float(ent_h.double().sum())
    ent, rel = ent_h.to(device), rel_h.to(device)
    del ent_h, rel_h
    ctx = runtime.Context(kind, ent, rel, norm=2, conve=conve, device=local)
    for kv in args.opt:
        name, value = kv.split("=")
        ctx.set_option(name, int(value))
    hp = runtime.make_hp(kind, cfg["hp"])

    # ONE batch, cut into `world` contiguous slices of near-equal cost; every rank adds the base mimic (job 0)
    batch = make_jobs(cfg, D, C)
    n_cand = len(batch["jobs"]) - 1
    costs = [len(f) for f in batch["jobs"][1:]]
    bounds = parallel.shard_bounds(costs, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    my_jobs = [0] + list(range(1 + lo, 1 + hi))
    arrs, triples, flt_off, flt_ids = build_arrays(cfg, batch, my_jobs)
    mode = runtime.RANK_ENGINE_MIN if kind == "TransE" else runtime.RANK_ENGINE_MAX
    work = algorithmic_work(cfg, D, arrs)
    dtypes = dict(init_rows=torch.float32, row_off=torch.int64, rows_per_epoch=torch.int32, pos=torch.int32,
                  neg=torch.int32, pos_off=torch.int64, pos_ids=torch.int32, fact_off=torch.int64, facts=torch.int32,
                  pos_idx=torch.uint16, neg_code=torch.int32)
    host = {k: v for k, v in arrs.items() if k != "static_epochs" and v is not None}
    max_rows, total_rows = int(arrs["rows_per_epoch"].max()), int(arrs["row_off"][-1])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device) if ent.numel() * 4 < (200 << 20) else None
    n_parity = min(8, hi - lo)

    # End-to-end pipeline: the step's inputs live in pinned host memory (plans assembles the batch arrays there) and
    # travel to one of TWO sets of device buffers on a copy stream, so the H2D of step k+1 runs behind the kernels of
    # step k (the C ABI takes the stream every call runs on; runtime.Context passes torch's current stream).
    def pinned(a, dt):
        t = torch.from_numpy(np.ascontiguousarray(a))
        t = t.to(dt) if t.dtype != dt else t
        return t if t.is_pinned() else t.pin_memory()

    pin = {k: pinned(v, dtypes[k]) for k, v in host.items()}
    pin["_triples"], pin["_flt_off"], pin["_flt_ids"] = pinned(triples, torch.int32), pinned(flt_off, torch.int64), pinned(flt_ids, torch.int32)
    h2d = sum(t.numel() * t.element_size() for t in pin.values())
    bufs = [{k: torch.empty_like(t, device=device) for k, t in pin.items()} for _ in range(2)]
    main_stream, copy_stream = torch.cuda.current_stream(device), torch.cuda.Stream(device=device)
    ready = [torch.cuda.Event() for _ in range(2)]  # inputs of the slot have arrived
    done = [torch.cuda.Event() for _ in range(2)]   # the kernels that read the slot have finished

    copy_events = []

    def issue_copy(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[slot])
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(copy_stream)
            for k, t in pin.items():
                bufs[slot][k].copy_(t, non_blocking=True)
            c1.record(copy_stream)
            ready[slot].record(copy_stream)
            copy_events.append((c0, c1))

    step_no = [0]
    out_pin = [torch.empty((1 + n_cand, 2), dtype=torch.float32, pin_memory=True) for _ in range(2)]

    def step(keep_rows=False):
        slot = step_no[0] & 1
        step_no[0] += 1
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        main_stream.wait_event(ready[slot])
        dev = bufs[slot]
        ev[0].record()
        rows = ctx.post_train(hp, static_epochs=arrs["static_epochs"], max_rows_per_epoch=max_rows, total_rows=total_rows,
                              **{k: v for k, v in dev.items() if not k.startswith("_")})
        ts, bs, rk = ctx.filtered_rank(dev["_triples"], mode, mimic_rows=rows, flt_off=dev["_flt_off"], flt_ids=dev["_flt_ids"])
        ev[1].record()
        done[slot].record(main_stream)
        issue_copy(slot ^ 1)  # the next step's inputs stream in behind this step's kernels
        pair = torch.stack([ts, rk.to(torch.float32)], 1)  # [1 + n_local, 2]; ranks <= N + 1 < 2^24 are exact in fp32
        res = torch.cat([pair[:1], parallel.gather_results(pair[1:], bounds)], 0)  # one all-gather when world > 1
        out_pin[slot].copy_(res, non_blocking=True)  # D2H of the step's result into pinned memory, read after the sync
        return ev, out_pin[slot], (rows[1:1 + n_parity].cpu() if keep_rows else None)

    def relevance(out):  # post_training_engine.py:136-145 on the host (C floats)
        sc, rk = out[:, 0].double().numpy(), out[:, 1].double().numpy()
        d = sc[1:] - sc[:1] if kind == "TransE" else sc[:1] - sc[1:]
        return (rk[1:] - rk[:1]) + 1.0 / (1.0 + np.exp(-d))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def agree(flag):  # every rank takes rank 0's decision
        if world == 1:
            return bool(flag)
        t = torch.tensor([1 if flag else 0], device=device)
        dist.broadcast(t, 0)
        return bool(t.item())

    # Wall budget: the driver kills a run at its per-N limit, which loses the whole record.  The timed loop stops early
    # (and the line carries the number of steps actually timed) rather than run into it.
    budget = float(os.environ.get("KP_BENCH_BUDGET_S", 800.0))
    elapsed = lambda: time.time() - T_PROCESS_START

    # nvidia-smi starts before the warm-up: its start-up (fork, NVML initialisation) must not fall into a short timed region
    sampler = ClockSampler(local) if rank == 0 else None
    issue_copy(0)
    warm_done, step_s = 0, None
    for i in range(args.warmup):
        # at least 3 warm-up steps (clocks and caches settle within the first: a step is >= tens of ms of dense work);
        # beyond that only while the remaining warm-up AND every requested timed step still fit the budget
        if agree(i >= 3 and step_s is not None and elapsed() + (args.warmup - i + args.steps) * step_s * 1.02 + 15.0 > budget):
            log(f"[rank {rank}] wall budget {budget:.0f} s: {i} warm-up steps instead of {args.warmup} so that the timed steps fit")
            break
        t = time.time()
        step()
        torch.cuda.synchronize()
        if flush is not None:
            flush.fill_(i)
        step_s = time.time() - t
        warm_done += 1
        log(f"[rank {rank}] warm-up step {i}: {step_s:.2f} s")

    ctx.set_option("timing", 1)
    ctx.stat("reset")
    barrier()
    launches0, w0 = ctx.launches, time.time()
    t_kernel = 0.0
    steps_done, rows_p = 0, None
    e2e_start, e2e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    events = []
    e2e_start.record()
    for i in range(args.steps):
        if agree(steps_done >= 1 and step_s is not None and elapsed() + 1.1 * step_s + 15.0 > budget):
            log(f"[rank {rank}] wall budget {budget:.0f} s: stopping after {steps_done} of {args.steps} timed steps")
            break
        t = time.time()
        ev, out, rp = step(keep_rows=(rank == 0 and i == 0))
        rows_p = rp if rp is not None else rows_p
        events.append(ev)
        steps_done += 1
        if flush is not None:
            flush.fill_(i)  # L2 flush between timed steps (0.04 ms of the end-to-end window, outside the kernel brackets)
        if step_s is None or step_s > 0.05:
            torch.cuda.synchronize()  # long steps: the host stays in step with the device (the wall budget needs real times);
            step_s = time.time() - t  # short steps: it runs ahead through the two buffer sets, results land in pinned memory
    e2e_end.record()
    barrier()
    t_kernel = sum(a.elapsed_time(b) for a, b in events)
    # end to end: every timed step's H2D (one per step, overlapped), kernels, all-gather and D2H, first launch to last read
    t_e2e = e2e_start.elapsed_time(e2e_end)
    w1 = time.time()
    launches = ctx.launches - launches0
    rel_vals = relevance(out)
    dom_ms, dom_n = ctx.stat("ms_" + work["what"]), ctx.stat("n_" + work["what"])
    breakdown = {c: round(ctx.stat("ms_" + c) / max(steps_done, 1), 3) for c in ("pass", "flash", "transe_train", "update", "conv")}
    ctx.set_option("timing", 0)

    t = torch.tensor([t_kernel, t_e2e], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_kernel, t_e2e = t.tolist()

    if rank == 0:
        total = C * steps_done  # ONE batch of C candidates per step, whatever the number of GPUs
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        achieved = peak = unit = src = None
        if work["bound"] == "tensor":
            peak, unit, src = peaks.get("bf16_tflops_sustained", 1400.0), "TFLOP/s", "bf16_tflops_sustained"
            achieved = work["units"] / (dom_ms / dom_n * 1e-3) / 1e12 if dom_n else None
        else:
            peak, unit, src = peaks.get("hbm_gbs", 6650.0), "GB/s", "hbm_gbs"
            achieved = work["units"] / (dom_ms / dom_n * 1e-3) / 1e9 if dom_n else None
        src += " of measured (MEASURED_PEAKS.json)" if peaks else " of fallback"
        traffic = None  # DRAM bytes per launch of the dominant kernel: from the committed ncu capture of this exact size
        for name in ("r02_traffic.json", "r01_traffic.json"):
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", name)))[f"{args.workload}/{C}/{world}"]["dram_bytes_per_launch"]
                break
            except Exception:
                pass
        line = {
            "metric": "candidate explanations evaluated/sec (post-train + filtered rank)",
            "value": total / (t_kernel * 1e-3), "unit": "candidates/s", "n_gpus": world, "steps": steps_done,
            "warmup": warm_done, "ms_per_step": t_kernel / steps_done, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 (bf16x3 split products, fp32 accumulate)" if work["bound"] == "tensor" else "f32",
            "data": "synthetic",
            "config": {"workload": args.workload, "model": kind, "entities": N, "row_floats": D, "relations": cfg["R"],
                       "candidates_per_step": C, "post_trainings_per_step": n_cand + world,
                       "candidates_per_gpu": [bounds[r + 1] - bounds[r] for r in range(world)],
                       "facts_per_candidate": list(cfg["T"]), "epochs": cfg["hp"]["epochs"],
                       "parallelism": f"one {C}-candidate batch cut into {world} cost-balanced slices (parallel.shard_bounds), tables "
                                      f"replicated, base mimic on every rank, one all-gather of (score, rank)",
                       "l2": "tables exceed the 126 MB L2" if flush is None else "256 MB L2 flush between timed steps"},
            "e2e": {"value": total / (t_e2e * 1e-3), "unit": "candidates/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(out.numel() * 4), "ms_per_step": t_e2e / steps_done,
                    "h2d_ms_per_step": float(np.median([a.elapsed_time(b) for a, b in copy_events[-max(steps_done, 1):]])),
                    "pipeline": "pinned host arrays -> two device buffer sets on a copy stream: the H2D of step k+1 overlaps the kernels of step k"},
            "gpu_launches": int(launches),
            "roofline": {"bound": work["bound"], "kernel": work["what"], "achieved": achieved, "peak": peak, "unit": unit,
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": src,
                         "launches_timed": int(dom_n), "avg_launch_ms": (dom_ms / dom_n) if dom_n else None,
                         "share_of_step": dom_ms / t_kernel if t_kernel else None,
                         "scope": "rank 0's slice" if world > 1 else "the whole batch",
                         # fp32 parity on bf16 tensor cores: every product is hi*hi + hi*lo + lo*hi, so the
                         # tensor pipe executes 3 MMAs per algorithmic one and `frac` cannot exceed 1/3
                         **({"mma_per_product": 3, "executed_frac": 3 * achieved / peak if achieved else None}
                            if work["bound"] == "tensor" else {}),
                         **({"note": "latency / issue bound kernel (DRAM idle in ncu): the HBM figure is the gather rate, not a roofline claim"}
                            if work["bound"] == "latency" else {})},
            "kernel_ms_per_step": breakdown,  # CUDA-event time of the library's kernels by category (rank 0)
            "clocks": sampler.summary(w0, w1) if sampler else None,
            "relevance_checksum": float(np.nansum(rel_vals)),
        }
        if steps_done != args.steps or warm_done != args.warmup:
            line["steps_requested"], line["warmup_requested"] = args.steps, args.warmup
            line["stopped_early"] = f"wall budget {budget:.0f} s (KP_BENCH_BUDGET_S)"
        if leg is not None:
            cb, parity = collect_cpu_leg(leg, leg_path, cfg, conv, ent_checksum, out, rows_p)
            line["cpu_baseline"] = cb
            if parity is not None:
                line["parity"] = parity
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def collect_cpu_leg(leg, path, cfg, conv, ent_checksum, out, rows_p):
    """Wait (bounded) for the child's last candidate, then compare its rows / scores / ranks with the GPU's."""
    try:
        leg.wait(timeout=float(os.environ.get("KP_CPU_LEG_WAIT_S", 120.0)))
    except subprocess.TimeoutExpired:
        leg.kill()  # the child we started, by its own handle
        leg.wait()
    try:
        r = json.load(open(path))
    except Exception:
        return {"value": None, "unit": "candidates/s", "cores": os.cpu_count(), "kind": "port", "sample": "the CPU leg produced no result"}, None
    n, t_total = len(r["times"]), float(sum(r["times"]))
    cb = {"value": n / t_total / conv, "unit": "candidates/s", "cores": r["cores"], "kind": "port",
          "sample": f"the first {n} candidates of the same batch (post-training, all epochs, + filtered rank), {t_total:.1f} s of CPU time, "
                    f"run by a child process while the GPU warmed up" + (f"; one candidate = {conv} of them" if conv > 1 else "")}
    parity = None
    if any(float(x) != 0.0 for x in cfg.get("dropout", ())):
        parity = {"skipped": "dropout > 0: the reference draws its masks from torch's generator, the device kernels from a counter-based "
                             "one (DESIGN.md section 7) -- rows are comparable at rate 0 only (tests/test_gpu_parity.py, test_gpu_dbpedia50.py)"}
    elif rows_p is not None and abs(r["ent_checksum"] - ent_checksum) <= 1e-9 * max(1.0, abs(ent_checksum)):
        jobs = r["jobs"]
        g_sc = np.array([float(out[j, 0]) for j in jobs], dtype=np.float64)
        g_rk = [int(out[j, 1]) for j in jobs]
        c_sc = np.array(r["scores"], dtype=np.float64)
        parity = {"candidates": n, "ranks_equal": g_rk == [int(x) for x in r["ranks"]],
                  "max_rel_err_scores": float(np.max(np.abs(g_sc - c_sc) / np.maximum(np.abs(c_sc), 1e-30)))}
        parity["ranks"] = {"gpu": g_rk, "oracle": [int(x) for x in r["ranks"]]}
        if all(x is not None for x in r["rows"]) and max(jobs) <= rows_p.shape[0]:
            c_rows = np.array(r["rows"], dtype=np.float64)
            g_rows = np.stack([rows_p[j - 1].double().numpy() for j in jobs])
            parity["max_rel_err_rows"] = float(np.abs(g_rows - c_rows).max() / np.abs(c_rows).max())
            if r.get("row64") is not None:
                r64 = np.array(r["row64"], dtype=np.float64)
                parity["rows_vs_fp64_oracle"] = float(np.abs(g_rows[0] - r64).max() / np.abs(r64).max())
                parity["oracle_fp32_vs_fp64"] = float(np.abs(c_rows[0] - r64).max() / np.abs(r64).max())
                parity["note"] = ("Adagrad's update lr * g / sqrt(sum g^2) is scale-invariant: components with a near-zero gradient turn fp32 "
                                  "summation-order differences over all entities into row differences, so two fp32 implementations agree "
                                  "only to `oracle_fp32_vs_fp64` here (tests/test_gpu_full_size.py); ranks among ~1e6 near-equal scores move with them")
        parity["against"] = "oracle port (oracle/kelpie_oracle.py) on the same tables, init rows, facts and filters"
    elif rows_p is not None:
        parity = {"skipped": "the child generated different tables (checksum mismatch)"}
    return cb, parity


if __name__ == "__main__":
    main()
