"""Host-side id-triple dataset with the attributes the hot path reads from the reference's
`Dataset` (src/data/dataset.py:97-141, 282-352).  Label handling, PyKEEN loading and the
semantic (class) tables stay with the reference: they are out of scope (SURVEY.md section 2
row 7); a reference `Dataset` object can be passed to the engines instead of this class.
"""
import os
from collections import defaultdict

import numpy as np

from .names import MANY_TO_MANY, MANY_TO_ONE, ONE_TO_MANY, ONE_TO_ONE


class Dataset:
    def __init__(self, name, train, valid, test, num_entities, num_relations):
        self.name = name
        self._train = np.asarray(train, dtype=np.int64).reshape(-1, 3)
        self._valid = np.asarray(valid, dtype=np.int64).reshape(-1, 3)
        self._test = np.asarray(test, dtype=np.int64).reshape(-1, 3)
        self._num_entities = int(num_entities)
        self._num_relations = int(num_relations)
        self.id_to_entity = {i: f"e{i}" for i in range(self._num_entities)}
        self.id_to_relation = {i: f"r{i}" for i in range(self._num_relations)}
        self.entity_to_id = {v: k for k, v in self.id_to_entity.items()}
        self.relation_to_id = {v: k for k, v in self.id_to_relation.items()}
        self._index()

    # -- construction helpers ---------------------------------------------------------
    @classmethod
    def from_npz(cls, path, name=None):
        z = np.load(path)
        return cls(name or os.path.basename(path), z["train"], z["valid"], z["test"], int(z["n_ent"]), int(z["n_rel"]))

    @classmethod
    def from_tsv(cls, directory, name=None):
        """train.txt / valid.txt / test.txt of tab-separated labels; ids = sorted training labels,
        valid/test rows with unseen labels dropped (what the reference gets from PyKEEN)."""

        def read(fn):
            with open(os.path.join(directory, fn), encoding="utf-8") as f:
                return [ln.rstrip("\n").split("\t") for ln in f if ln.count("\t") == 2]

        tr, va, te = read("train.txt"), read("valid.txt"), read("test.txt")
        ents = sorted({h for h, _, _ in tr} | {t for _, _, t in tr})
        rels = sorted({r for _, r, _ in tr})
        e2i, r2i = {e: i for i, e in enumerate(ents)}, {r: i for i, r in enumerate(rels)}
        ids = lambda rows: [(e2i[h], r2i[r], e2i[t]) for h, r, t in rows if h in e2i and t in e2i and r in r2i]
        ds = cls(name or os.path.basename(directory), ids(tr), ids(va), ids(te), len(ents), len(rels))
        ds.entity_to_id, ds.relation_to_id = e2i, r2i
        ds.id_to_entity = {v: k for k, v in e2i.items()}
        ds.id_to_relation = {v: k for k, v in r2i.items()}
        return ds

    def _index(self):
        R = self._num_relations
        self.entity_to_training_triples = defaultdict(list)
        self.entity_to_validation_triples = defaultdict(list)
        self.entity_to_testing_triples = defaultdict(list)
        for src, dst in ((self._train, self.entity_to_training_triples),
                         (self._valid, self.entity_to_validation_triples),
                         (self._test, self.entity_to_testing_triples)):
            for s, p, o in src:
                dst[s].append((s, p, o))
                dst[o].append((s, p, o))
        for e in list(self.entity_to_training_triples):  # dataset.py:113-116
            self.entity_to_training_triples[e] = list(set(self.entity_to_training_triples[e]))
        self.entity_to_degree = {e: len(self.entity_to_training_triples[e])
                                 for e in set(self.entity_to_training_triples) | set(self.entity_to_validation_triples)
                                 | set(self.entity_to_testing_triples)}
        self.train_to_filter = defaultdict(list)
        for s, p, o in self._train:
            self.train_to_filter[(s, p)].append(o)
            self.train_to_filter[(o, p + R)].append(s)
        self.to_filter = defaultdict(list)
        for s, p, o in self.all_triples:
            self.to_filter[(s, p)].append(o)
            self.to_filter[(o, p + R)].append(s)
        # the same multiset as rows (entity, relation, id) for the device CSR builder (kp_filter_build); edits are logged
        a = np.asarray(self.all_triples, dtype=np.int64).reshape(-1, 3)
        self._filter_base = np.vstack((a, np.stack((a[:, 2], a[:, 1] + R, a[:, 0]), 1))).astype(np.int32)
        self._filter_added, self._filter_removed = [], []
        self._compute_relation_to_type()

    # -- attributes of the reference class ------------------------------------------------
    @property
    def training_triples(self):
        return self._train

    @property
    def validation_triples(self):
        return self._valid

    @property
    def testing_triples(self):
        return self._test

    @property
    def all_triples(self):
        return np.vstack([self._train, self._valid, self._test])

    @property
    def num_entities(self):
        return self._num_entities

    @property
    def num_relations(self):
        return self._num_relations

    def labels_triple(self, t):
        s, p, o = t
        return (self.id_to_entity[s], self.id_to_relation[p], self.id_to_entity[o])

    def labels_triples(self, ts):
        return [self.labels_triple(t) for t in ts]

    def ids_triple(self, t):
        s, p, o = t
        return (self.entity_to_id[s], self.relation_to_id[p], self.entity_to_id[o])

    def printable_triple(self, t):
        s, p, o = self.labels_triple(t)
        return f"<{s}, {p}, {o}>"

    def printable_nple(self, nple):
        return " +\n\t\t".join(self.printable_triple(t) for t in nple)

    def _compute_relation_to_type(self):
        """dataset.py:282-317."""
        R = self._num_relations
        s_num, o_num = defaultdict(list), defaultdict(list)
        for (e, r) in self.train_to_filter:
            n = len(self.to_filter[(e, r)])
            (s_num[r - R] if r >= R else o_num[r]).append(n)
        self.relation_to_type = {}
        for r in s_num:
            a, b = np.average(s_num[r]), np.average(o_num[r])
            if a > 1.2 and b > 1.2:
                self.relation_to_type[r] = MANY_TO_MANY
            elif a > 1.2:
                self.relation_to_type[r] = MANY_TO_ONE
            elif b > 1.2:
                self.relation_to_type[r] = ONE_TO_MANY
            else:
                self.relation_to_type[r] = ONE_TO_ONE

    # -- in-place edits used by verify_explanations (dataset.py:242-280) --------------------------
    def add_training_triple(self, triple):
        s, p, o = (int(x) for x in triple)
        self._train = np.vstack((self._train, np.array([[s, p, o]], dtype=np.int64)))
        self.entity_to_training_triples[s].append((s, p, o))
        self.entity_to_training_triples[o].append((s, p, o))
        self.entity_to_degree[s] = self.entity_to_degree.get(s, 0) + 1
        self.entity_to_degree[o] = self.entity_to_degree.get(o, 0) + 1
        self.to_filter[(s, p)].append(o)  # like the reference, only the direct key is maintained
        self.train_to_filter[(s, p)].append(o)
        self._filter_added.append((s, p, o))

    def add_training_triples(self, triples):
        for t in triples:
            self.add_training_triple(t)

    def remove_training_triple(self, triple):
        s, p, o = (int(x) for x in triple)
        t = self._train
        self._train = t[~((t[:, 0] == s) & (t[:, 1] == p) & (t[:, 2] == o))]
        self.entity_to_training_triples[s].remove((s, p, o))
        if s != o:
            self.entity_to_training_triples[o].remove((s, p, o))
        self.entity_to_degree[s] -= 1
        if s != o:
            self.entity_to_degree[o] -= 1
        self.to_filter[(s, p)].remove(o)
        self.train_to_filter[(s, p)].remove(o)
        self._filter_removed.append((s, p, o))

    def remove_training_triples(self, triples):
        for t in set(tuple(int(x) for x in t) for t in triples):
            self.remove_training_triple(t)

    def filter_facts(self):
        """The multiset `to_filter` as int32 rows (entity, relation, id) -- input of the device CSR builder
        (runtime.Context.build_filter / kp_filter_build).  Edits follow the dict: an added training fact contributes its
        DIRECT key only, a removed one takes away ONE occurrence of its direct row (dataset.py:242-280)."""
        rows = self._filter_base
        if self._filter_added:
            rows = np.vstack((rows, np.asarray(self._filter_added, dtype=np.int32)))
        if self._filter_removed:
            keep = np.ones(len(rows), dtype=bool)
            for s, p, o in self._filter_removed:
                hit = np.flatnonzero(keep & (rows[:, 0] == s) & (rows[:, 1] == p) & (rows[:, 2] == o))
                keep[hit[0]] = False  # list.remove: first occurrence; which copy goes does not matter to a multiset
            rows = rows[keep]
        return np.ascontiguousarray(rows, dtype=np.int32)

    def invert_triples(self, triples):
        """dataset.py:319-331."""
        t = np.asarray(triples)
        out = np.copy(t)
        out[:, 0] = t[:, 2]
        out[:, 2] = t[:, 0]
        out[:, 1] = t[:, 1] + self._num_relations
        return out

    @staticmethod
    def replace_entity_in_triple(triple, old_entity, new_entity):
        s, p, o = triple
        return (new_entity if s == old_entity else s, p, new_entity if o == old_entity else o)

    @staticmethod
    def replace_entity_in_triples(triples, old_entity, new_entity):
        return [Dataset.replace_entity_in_triple(t, old_entity, new_entity) for t in triples]
