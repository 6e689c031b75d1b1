"""`RelevanceEngine` (src/relevance_engines/engine.py:13-126) with the all-entity scoring of
every eligible head done by ONE filtered-rank launch instead of batches of 4 + D2H."""
import random
from collections import defaultdict

import numpy as np

from .. import runtime
from ..data.names import MANY_TO_ONE, ONE_TO_ONE
from ..link_prediction.models.model import context_for


class RelevanceEngine:
    def __init__(self, model, dataset):
        self.model = model
        self.dataset = dataset
        self.o_to_training_triples = defaultdict(list)
        for h, r, t in dataset.training_triples:
            self.o_to_training_triples[t].append((h, r, t))

    def convertible_entities(self, pred, degree_cap=None, criage=False):
        """engine.py:62-124: eligible heads e whose (e,p,o) is not already filtered-rank 1."""
        s, p, o = pred
        ds = self.dataset
        # engine.py:70-88 (host eligibility rules), one pass with the lookups hoisted out of the loop
        degree, to_filter = ds.entity_to_degree, ds.to_filter
        functional = ds.relation_to_type[p] in [ONE_TO_ONE, MANY_TO_ONE]
        cap = degree_cap if degree_cap else float("inf")
        heads = self.o_to_training_triples if criage else None
        entities = []
        for entity in range(ds.num_entities):
            if entity == s or not (1 <= degree[entity] <= cap):
                continue
            if heads is not None and entity not in heads:
                continue
            key = (entity, p)
            if key in to_filter and (functional or o in to_filter[key]):
                continue
            entities.append(entity)
        if len(entities) == 0:
            return []
        triples = np.empty((len(entities), 3), dtype=np.int64)
        triples[:, 0], triples[:, 1], triples[:, 2] = entities, p, o
        ctx = context_for(self.model)
        mode = runtime.RANK_MODEL
        ts, _, _, cnt = ctx.filtered_rank(triples, mode, counters=True)
        ts, cnt = ts.cpu().numpy(), cnt.cpu().numpy()
        # engine.py:113-120: 1e6 > target > min(scores)  /  -1e6 < target < max(scores)
        bound = (ts < 1e6) if self.model.is_minimizer() else (ts > -1e6)
        keep = (cnt[:, 0] > 0) & bound
        return [e for e, k in zip(entities, keep) if k]

    def select_entities_to_convert(self, pred, k, degree_cap=None, criage=False):
        overall = self.convertible_entities(pred, degree_cap, criage)
        if len(overall) == 0 and not hasattr(self, "entities_to_convert"):
            self.entities_to_convert = []
        entities = random.sample(overall, k=min(k, len(overall)))  # engine.py:125
        self.entities_to_convert = entities
