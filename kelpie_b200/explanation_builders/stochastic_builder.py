"""Batched emitter of candidate explanations with the control flow of the reference's
`StochasticBuilder` (src/explanation_builders/stochastic_builder.py:13-192).

The reference evaluates one candidate at a time (`engine.compute_relevance`, :121,:150) and
decides after each one whether to stop (threshold xsi :159, sliding-window stochastic early
stop :163-173).  Here candidates are emitted to the engine in BATCHES (`compute_relevances`)
and the stop rule is replayed over the batch afterwards:

  * the engine draws every candidate's random numbers in sequential-call order and records a
    snapshot of the torch / numpy generators after each candidate;
  * the builder's own `random.random()` draws (:164) are made while replaying the decisions,
    exactly where the sequential loop makes them (a different generator, so the order relative
    to the engine's draws is immaterial);
  * if the loop stops inside a batch, the generators are rewound to the snapshot taken after the
    last candidate the sequential loop would have evaluated, and the speculative results
    beyond it are discarded.

Hence the relevances, the number of evaluated candidates (`#relevances`) and the selected
explanations are those of the sequential loop.  Summarisation is host-side graph code outside
this package: any object with `summarize(entity, triples)` / `map_rule(rule)` (e.g. the
reference's Simulation / Bisimulation) can be passed in.
"""
import inspect
import itertools
import random
import time

import numpy as np
import torch

from .. import key, plans


def _snapshot():
    state = [torch.get_rng_state(), np.random.get_state()]
    if torch.cuda.is_available() and torch.cuda.is_initialized():
        state.append(torch.cuda.get_rng_state())
    return state


def _restore(state):
    torch.set_rng_state(state[0])
    if isinstance(state[1], bytes):  # the engine's cheap copy of the generator words (plans.HostReplay.numpy_snapshot)
        plans.HostReplay.numpy_restore(state[1])
    else:
        np.random.set_state(state[1])
    if len(state) > 2:
        torch.cuda.set_rng_state(state[2])


class StochasticBuilder:
    def __init__(self, xsi, engine, summarization=None, max_explanation_length: int = 4, batch_size: int = 32):
        self.dataset = engine.dataset
        self.length_cap = max_explanation_length
        self.window_size = 10
        self.xsi = xsi
        self.engine = engine
        self.summarization = summarization  # object with summarize()/map_rule(), or None
        self.batch_size = batch_size        # candidates emitted per engine call in the compound phase
        self._batched = None

    # -- engine access ----------------------------------------------------------------------
    def _relevances(self, pred, rules):
        """[rule] -> ([relevance], [generator snapshot after each rule])."""
        if self._batched is None:  # looked up once per builder: does the engine take whole batches and return snapshots?
            batched = getattr(self.engine, "compute_relevances", None)
            self._batched = batched is not None and "snapshots" in inspect.signature(batched).parameters
        if self._batched:
            return self.engine.compute_relevances(pred, rules, snapshots=True)  # errors raised inside the engine propagate
        rels, snaps = [], []
        for r in rules:  # engines without a batch entry point (e.g. the reference's own)
            rels.append(self.engine.compute_relevance(pred, r))
            snaps.append(_snapshot())
        return rels, snaps

    def _map(self, rule):
        return self.summarization.map_rule(rule) if self.summarization else list(rule)

    # -- stochastic_builder.py:33-107 -----------------------------------------------------------
    def build_explanations(self, pred, candidate_triples: list, k: int = 10):
        start = time.time()
        pred_head = pred[0]
        if self.summarization is not None:
            summary = self.summarization.summarize(pred_head, candidate_triples)
            if len(summary) > 0:
                candidate_triples = summary
            else:
                self.summarization = None

        triple_to_rel = self.explore_singleton_rules(pred, candidate_triples)
        rule_to_rel = [((t,), rel) for (t, rel) in sorted(triple_to_rel.items(), key=key, reverse=True)]
        triples_number = len(triple_to_rel)
        rels_num = triples_number
        _, best = rule_to_rel[0]
        if not best > self.xsi:
            for rule_length in range(2, min(triples_number, self.length_cap) + 1):
                cur, cur_num = self.explore_compound_rules(pred, candidate_triples, rule_length, triple_to_rel)
                rels_num += cur_num
                cur = sorted(cur.items(), key=key, reverse=True)
                rule_to_rel += cur
                _, current_best = cur[0]
                if current_best > best:
                    best = current_best
                if best > self.xsi:
                    break

        rule_to_rel = sorted(rule_to_rel, key=lambda x: (x[1], 1 / len(x[0])), reverse=True)[:k]
        if self.summarization:
            mapped = []
            for rule, rel in rule_to_rel:
                mapped_rule = self.dataset.labels_triples(self.summarization.map_rule(rule))
                labels_rule = [([self.dataset.id_to_entity[e] for e in s_part], self.dataset.id_to_relation[p],
                                [self.dataset.id_to_entity[e] for e in o_part]) for s_part, p, o_part in rule]
                mapped.append((labels_rule, mapped_rule, rel))
        else:
            mapped = [(self.dataset.labels_triples(rule), rel) for rule, rel in rule_to_rel]
        return {"triple": self.dataset.labels_triple(pred), "rule_to_relevance": mapped, "#relevances": rels_num,
                "execution_time": time.time() - start}

    # -- :110-124 : every singleton is evaluated, so one batch reproduces the loop exactly --------
    def explore_singleton_rules(self, pred, triples: list):
        rules = [self._map([t]) for t in triples]
        rels, _ = self._relevances(pred, rules)
        return {t: r for t, r in zip(triples, rels)}

    # -- :126-175 ------------------------------------------------------------------------------
    def explore_compound_rules(self, pred, triples: list, length: int, triple_to_relevance: dict):
        rules = [(r, self.compute_rule_prescore(r, triple_to_relevance)) for r in itertools.combinations(triples, length)]
        rules = sorted(rules, key=lambda x: x[1], reverse=True)

        terminate = False
        best = -1e6
        sliding_window = [None for _ in range(self.window_size)]
        rule_to_relevance = {}
        computed = 0
        i = 0
        while i < len(rules) and not terminate:
            chunk = rules[i:i + self.batch_size]
            before = _snapshot()
            rels, snaps = self._relevances(pred, [self._map(r) for r, _ in chunk])
            consumed = 0
            done = False
            for j, ((rule, _), relevance) in enumerate(zip(chunk, rels)):
                idx = i + j
                rule_to_relevance[rule] = relevance
                computed += 1
                consumed = j + 1
                sliding_window[idx % self.window_size] = relevance
                if relevance > self.xsi:  # :159-160
                    done = True
                    break
                elif relevance >= best:
                    best = relevance
                elif idx >= self.window_size:  # :163-165
                    avg_window_relevance = sum(sliding_window) / self.window_size
                    terminate_threshold = avg_window_relevance / best
                    terminate = random.random() > terminate_threshold
                    if terminate:
                        break
            if consumed < len(chunk):  # stopped inside the batch: rewind the speculative draws
                _restore(snaps[consumed - 1] if consumed > 0 else before)
            if done:
                return rule_to_relevance, computed
            i += len(chunk)
        return rule_to_relevance, computed

    def compute_rule_prescore(self, rule, triple_to_relevance):
        return sum(triple_to_relevance[t] for t in rule)
