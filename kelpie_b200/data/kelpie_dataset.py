"""Mimic overlay of one explained entity (src/data/kelpie_dataset.py:10-203).

Same attributes and methods as the reference class, but copy-on-write: only the filter
lists whose key or content mentions the mimic are copied (the reference deep-copies the
whole `to_filter` / `train_to_filter` dicts per explained entity, seconds per prediction).
"""
import copy
from collections import ChainMap, defaultdict

import numpy as np

from .dataset import Dataset


class _Overlay(ChainMap):
    """dict view: touched keys live in maps[0], everything else falls through to the dataset."""

    def own(self, key):
        own = self.maps[0]
        if key not in own:
            own[key] = list(self.maps[1].get(key, []))
        return own[key]

    def __getitem__(self, key):
        try:
            return super().__getitem__(key)
        except KeyError:
            return self.own(key)  # defaultdict(list) behaviour of the reference


class KelpieDataset:
    def __init__(self, dataset, entity):
        self.dataset = dataset
        self.to_filter = _Overlay({}, dataset.to_filter)
        self.train_to_filter = _Overlay({}, dataset.train_to_filter)
        self.num_entities = dataset.num_entities + 1
        self.num_relations = dataset.num_relations
        self.original_entity = entity
        self.kelpie_entity = self.num_entities - 1
        rep = Dataset.replace_entity_in_triples
        self.kelpie_training_triples = rep(dataset.entity_to_training_triples[entity], entity, self.kelpie_entity)
        self.kelpie_validation_triples = rep(dataset.entity_to_validation_triples[entity], entity, self.kelpie_entity)
        self.kelpie_testing_triples = rep(dataset.entity_to_testing_triples[entity], entity, self.kelpie_entity)
        self.kelpie_training_triples_copy = list(self.kelpie_training_triples)  # rows are immutable tuples: as good as the reference's deepcopy
        R = self.num_relations
        for s, p, o in self.kelpie_training_triples:
            self.train_to_filter.own((s, p)).append(o)
            self.train_to_filter.own((o, p + R)).append(s)
        for s, p, o in self.kelpie_training_triples + self.kelpie_validation_triples + self.kelpie_testing_triples:
            self.to_filter.own((s, p)).append(o)
            self.to_filter.own((o, p + R)).append(s)
        self.kelpie_triple_to_index = {tuple(t): i for i, t in enumerate(self.kelpie_training_triples)}
        self.last_added_triples_number = 0
        self.last_removed_triples_number = 0
        self.last_filter_additions = defaultdict(list)
        self.last_filter_removals = defaultdict(list)

    def as_kelpie_triple(self, original_triple):
        if self.original_entity not in original_triple:
            raise Exception(f"Could not find the original entity {self.original_entity} in the passed triple {original_triple}")
        return Dataset.replace_entity_in_triple(original_triple, self.original_entity, self.kelpie_entity)

    def as_original_triple(self, kelpie_triple):
        if self.kelpie_entity not in kelpie_triple:
            raise Exception(f"Could not find the original entity {self.kelpie_entity} in the passed triple {kelpie_triple}")
        return Dataset.replace_entity_in_triple(kelpie_triple, self.kelpie_entity, self.original_entity)

    def _check(self, triples):
        for s, _, o in triples:
            assert self.original_entity == s or self.original_entity == o

    def add_training_triples(self, triples_to_add):
        """kelpie_dataset.py:98-128."""
        self._check(triples_to_add)
        R = self.num_relations
        self.last_added_triples_number = len(triples_to_add)
        self.last_filter_additions = defaultdict(list)
        new = Dataset.replace_entity_in_triples(triples_to_add, self.original_entity, self.kelpie_entity)
        for s, p, o in new:
            for flt in (self.to_filter, self.train_to_filter):
                flt.own((s, p)).append(o)
                flt.own((o, p + R)).append(s)
            self.last_filter_additions[(s, p)].append(o)
            self.last_filter_additions[(o, p + R)].append(s)
        self.kelpie_training_triples = list(self.kelpie_training_triples) + new

    def remove_training_triples(self, triples):
        """kelpie_dataset.py:130-158 (one occurrence removed per fact: multiset semantics)."""
        self._check(triples)
        R = self.num_relations
        self.last_removed_triples_number = len(triples)
        self.last_filter_removals = defaultdict(list)
        gone = Dataset.replace_entity_in_triples(triples, self.original_entity, self.kelpie_entity)
        for s, p, o in gone:
            for flt in (self.to_filter, self.train_to_filter):
                flt.own((s, p)).remove(o)
                flt.own((o, p + R)).remove(s)
            self.last_filter_removals[(s, p)].append(o)
            self.last_filter_removals[(o, p + R)].append(s)
        idx = {self.kelpie_triple_to_index[tuple(t)] for t in gone}
        self.kelpie_training_triples = [t for i, t in enumerate(self.kelpie_training_triples_copy) if i not in idx]

    def undo_removal(self):
        if self.last_removed_triples_number <= 0:
            raise Exception("No removal to undo.")
        self.kelpie_training_triples = list(self.kelpie_training_triples_copy)  # rows are immutable tuples
        for k, xs in self.last_filter_removals.items():
            for x in xs:
                self.to_filter.own(k).append(x)
                self.train_to_filter.own(k).append(x)
        self.last_removed_triples_number = 0
        self.last_filter_removals = defaultdict(list)

    def undo_addition(self):
        if self.last_added_triples_number <= 0:
            raise Exception("No addition to undo.")
        self.kelpie_training_triples = list(self.kelpie_training_triples_copy)
        for k, xs in self.last_filter_additions.items():
            for x in xs:
                self.to_filter.own(k).remove(x)
                self.train_to_filter.own(k).remove(x)
        self.last_added_triples_number = 0
        self.last_filter_additions = defaultdict(list)

    def invert_triples(self, triples):
        return self.dataset.invert_triples(triples)

    def printable_triple(self, triple):
        return self.dataset.printable_triple(triple)
