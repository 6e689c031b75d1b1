"""Necessary / sufficient post-training relevance engines
(src/relevance_engines/post_training_engine.py:17-207).

`compute_relevance(pred, triples)` keeps the reference's signature and result; the additive
`compute_relevances(pred, [rules])` evaluates a whole batch of candidate explanations with one
post-training launch sequence and one filtered-rank launch, drawing every random number in
the order the sequential reference loop would (so the selected explanations are identical).
"""
import math
from collections import OrderedDict

import numpy as np
import torch

from .. import plans, runtime
from ..data import Dataset, KelpieDataset
from ..link_prediction.models.model import KelpieModel, context_for
from ..link_prediction.models.conve import burn_conve_constructor_rng
from .engine import RelevanceEngine


class PostTrainingEngine(RelevanceEngine):
    @staticmethod
    def sigmoid(x):
        return 1 / (1 + math.exp(-x))

    def __init__(self, model, dataset, hp: dict):
        RelevanceEngine.__init__(self, model=model, dataset=dataset)
        self.hp = hp
        if isinstance(model, KelpieModel):
            raise Exception("Already a post-trainable KelpieModel.")
        # where KelpieTransE's xavier_normal_ draws (transe.py:93-95): the reference's mimic row
        # lives on the GPU, so "cuda"; the CPU-patched oracle draws on the CPU generator.
        self.rng_device = "cuda"
        # replay the CPU-generator draws of KelpieConvE's discarded layer initialisers
        self.replay_constructor_rng = True
        # None (default): shuffles / corruptions / permutations come from the reference's generators in the
        # reference's order, so the selected explanations are the reference's.  A numpy Generator here draws them
        # vectorised instead (same distributions, ~30x less host time per TransE candidate).
        self.fast_rng = None
        self.set_cache()

    def set_cache(self):
        self.base_pt_results = {}
        self.kelpie_dataset_cache_size = 20
        self.kelpie_dataset_cache = OrderedDict()

    def _get_kelpie_dataset(self, original_entity):
        if original_entity not in self.kelpie_dataset_cache:
            self.kelpie_dataset_cache[original_entity] = KelpieDataset(dataset=self.dataset, entity=original_entity)
            self.kelpie_dataset_cache.move_to_end(original_entity)
            if len(self.kelpie_dataset_cache) > self.kelpie_dataset_cache_size:
                self.kelpie_dataset_cache.popitem(last=False)
        return self.kelpie_dataset_cache[original_entity]

    # ---- mimic row initialisation (transe.py:92-95, complex.py:155-157, conve.py:202-209) ----
    def _init_row(self, init_tensor):
        name = self.model.name
        if name == "TransE":
            # xavier_normal_ of a [1, D] tensor = normal_(0, sqrt(2 / (fan_in + fan_out))) with fan_in = D, fan_out = 1, drawn
            # on the generator of the row's device; the values of init_tensor are overwritten, so they are not copied over
            rows, cols = init_tensor.shape
            row = torch.empty((rows, cols), dtype=init_tensor.dtype, device=self.rng_device)
            row.normal_(0.0, math.sqrt(2.0 / float(cols + rows)))
            return row if row.is_cuda else row.numpy()  # a device draw stays on the device: no D2H sync per candidate
        if name == "ComplEx":
            return (init_tensor.clone() * self.model.init_scale).numpy()
        if self.replay_constructor_rng:
            burn_conve_constructor_rng(self.model)
        return init_tensor.clone().numpy()

    def _apply(self, dataset, triples):
        raise NotImplementedError

    def _undo(self, dataset):
        raise NotImplementedError

    # ---- batched core ---------------------------------------------------------------------
    @staticmethod
    def _rng_snapshot():
        state = [torch.get_rng_state(), plans.HostReplay.numpy_snapshot() if plans.HostReplay.available() else np.random.get_state()]
        if torch.cuda.is_available() and torch.cuda.is_initialized():
            state.append(torch.cuda.get_rng_state())
        return state

    def individual_results(self, items, snapshots=None, owned=None):
        """[(pred, rule)] -> [(pt_results, base_pt_results)], RNG drawn in sequential-call order.

        One mimic post-training job per candidate plus one per prediction whose homologous
        (base) mimic is not cached yet (post_training_engine.py:46-62, 78-90).

        owned = (lo, hi) (candidate sharding over several GPUs, parallel.ShardedEngine): EVERY item's random numbers are
        drawn -- so the generators end where the single-process run leaves them -- but only items lo..hi-1 are
        post-trained; the result list then covers those items only.  Base mimics are computed on every rank."""
        model, kind = self.model, self.model.name
        ctx = context_for(model)
        N, R = self.dataset.num_entities, self.dataset.num_relations
        batch = plans.Batch(kind, N, R, self.hp, fast_rng=self.fast_rng)
        scratch = plans.Batch(kind, N, R, self.hp, fast_rng=self.fast_rng) if owned is not None else None
        job_triple, job_filter = [], []
        pending_base, slots = {}, []
        for index, (pred, rule) in enumerate(items):
            pred = tuple(int(x) for x in pred)
            ds = self._get_kelpie_dataset(pred[0])
            kp = ds.as_kelpie_triple(pred)
            init = torch.rand(1, model.dimension)  # post_training_engine.py:52 (CPU generator)
            base_row = self._init_row(init)  # :55 -- built (and drawn) even when the base is cached
            if pred not in self.base_pt_results and pred not in pending_base:
                pending_base[pred] = batch.add(ds.kelpie_training_triples, base_row)
                job_triple.append(kp)
                job_filter.append(sorted(set(ds.to_filter.get((kp[0], kp[1]), []))))
            pt_row = self._init_row(init)  # :59
            self._apply(ds, rule)
            if owned is None or owned[0] <= index < owned[1]:
                j = batch.add(ds.kelpie_training_triples, pt_row)
                job_triple.append(kp)
                job_filter.append(sorted(set(ds.to_filter.get((kp[0], kp[1]), []))))
                slots.append((pred, j))
            else:  # another rank's candidate: consume the generators exactly as its plan would
                scratch.add(ds.kelpie_training_triples, pt_row)
            self._undo(ds)
            if snapshots is not None:  # generator state after this candidate's draws (builder early stop)
                snapshots.append(self._rng_snapshot())

        if len(batch) == 0:  # an empty slice of a sharded batch whose base mimics are all cached
            return []
        arrs = batch.arrays(compact=True)  # TransE: 6-byte index rows over PCIe
        hp = runtime.make_hp(kind, self.hp)
        # ConvE dropout masks are counter-based (kp_dropout.cuh): seeded by the user's torch seed and a per-batch counter,
        # without touching the host generators (the reference draws its masks on the CUDA generator).  With a non-zero
        # dropout rate (DB100K config) the masks therefore depend on how candidates were batched -- relevances are then
        # reproducible for a given seed and batch size, not equal to a sequential run's; every rate is 0 in the DBpedia50
        # config, where the batched and the sequential results coincide.
        self._batches = getattr(self, "_batches", 0) + 1
        rows = ctx.post_train(hp, dropout_seed=((torch.initial_seed() & 0xFFFFFFFF) << 32) | self._batches, **arrs)
        flt_off = np.zeros(len(job_filter) + 1, dtype=np.int64)
        flt_off[1:] = np.cumsum([len(f) for f in job_filter])
        flt_ids = np.array([x for f in job_filter for x in f], dtype=np.int32)
        mode = runtime.RANK_ENGINE_MIN if model.is_minimizer() else runtime.RANK_ENGINE_MAX
        ts, bs, rk = ctx.filtered_rank(np.array(job_triple, dtype=np.int64), mode, mimic_rows=rows,
                                       flt_off=flt_off, flt_ids=flt_ids if len(flt_ids) else None)
        ts, bs, rk = ts.cpu().numpy(), bs.cpu().numpy(), rk.cpu().numpy()
        self.last_rows = rows

        def result(j):
            return {"target_score": float(ts[j]), "best_score": torch.tensor(float(bs[j])),
                    "target_rank": torch.tensor(int(rk[j])), "_rank": int(rk[j])}

        for pred, j in pending_base.items():
            self.base_pt_results[pred] = result(j)
        return [(result(j), self.base_pt_results[pred]) for pred, j in slots]

    def compute_relevance(self, pred, triples):
        [(pt, base)] = self.individual_results([(pred, triples)])
        return pt, base


class NecessaryPostTrainingEngine(PostTrainingEngine):
    def _apply(self, dataset, triples):
        dataset.remove_training_triples(triples)

    def _undo(self, dataset):
        dataset.undo_removal()

    def _relevance(self, pt, base):  # post_training_engine.py:136-145
        score_worsening = (pt["target_score"] - base["target_score"] if self.model.is_minimizer()
                           else base["target_score"] - pt["target_score"])
        if "_rank" in pt and "_rank" in base:  # int64 tensor + Python float = a float32 sum, without the tensor ops
            return float(np.float32(pt["_rank"] - base["_rank"]) + np.float32(self.sigmoid(score_worsening)))
        rank_worsening = pt["target_rank"] - base["target_rank"]
        return float(rank_worsening + self.sigmoid(score_worsening))

    def compute_relevance(self, pred, triples):
        return self.compute_relevances(pred, [triples])[0]

    def compute_relevances(self, pred, rules, snapshots=False, owned=None):
        """owned = (lo, hi): relevances of rules lo..hi-1 only (see individual_results); snapshots cover every rule."""
        snaps = [] if snapshots else None
        rels = [self._relevance(pt, base) for pt, base in self.individual_results([(pred, r) for r in rules], snaps, owned)]
        return (rels, snaps) if snapshots else rels


class SufficientPostTrainingEngine(PostTrainingEngine):
    def _apply(self, dataset, triples):
        dataset.add_training_triples(triples)

    def _undo(self, dataset):
        dataset.undo_addition()

    def _relevance(self, pt, base):  # post_training_engine.py:164-176
        score_improvement = (base["target_score"] - pt["target_score"] if self.model.is_minimizer()
                             else pt["target_score"] - base["target_score"])
        if "_rank" in pt and "_rank" in base:  # as above: the float32 sum of the reference's tensor expression
            relevance = float(np.float32(base["_rank"] - pt["_rank"]) + np.float32(self.sigmoid(score_improvement)))
            return relevance / float(base["_rank"])
        rank_improvement = base["target_rank"] - pt["target_rank"]
        relevance = float(rank_improvement + self.sigmoid(score_improvement))
        return relevance / float(base["target_rank"])

    def compute_individual_relevance(self, pred, triples):
        [(pt, base)] = self.individual_results([(pred, triples)])
        return self._relevance(pt, base)

    def compute_relevance(self, pred, rule):
        return self.compute_relevances(pred, [rule])[0]

    def compute_relevances(self, pred, rules, snapshots=False, owned=None):
        """post_training_engine.py:178-191 for every rule: rule-major, conversion-entity-minor.
        owned = (lo, hi): relevances of rules lo..hi-1 only (all their conversions stay on this rank)."""
        s = pred[0]
        items = []
        for rule in rules:
            for e in self.entities_to_convert:
                items.append((Dataset.replace_entity_in_triple(pred, s, e), Dataset.replace_entity_in_triples(rule, s, e)))
        snaps = [] if snapshots else None
        k = len(self.entities_to_convert)
        item_range = None if owned is None else (owned[0] * k, owned[1] * k)
        res = [self._relevance(pt, base) for pt, base in self.individual_results(items, snaps, item_range)]
        rels = [sum(res[i * k:(i + 1) * k]) / k for i in range(len(res) // k)]
        return (rels, snaps[k - 1::k]) if snapshots else rels
