"""StochasticBuilder batched emission == the sequential reference loop (stochastic_builder.py:33-175):
same relevances, same #relevances, same selected rules, same generator state afterwards.
Uses the MockEngine pattern of the reference's own builder test (test_stochastic_builder.py:7-11),
extended so that every relevance consumes torch / numpy random numbers like a real engine."""
import itertools
import random

import numpy as np
import pytest
import torch

from kelpie_b200.explanation_builders import StochasticBuilder


class FakeDataset:
    def labels_triple(self, t):
        return tuple(t)

    def labels_triples(self, ts):
        return [tuple(t) for t in ts]


class RngEngine:
    """relevance = f(rule) + noise drawn from torch AND numpy (so the draws of skipped candidates matter)."""

    def __init__(self, base, batch_api):
        self.dataset = FakeDataset()
        self.base, self.calls, self.batch_api = base, [], batch_api

    def compute_relevance(self, pred, rule):
        self.calls.append(tuple(rule))
        noise = float(torch.rand(1)) * 0.3 + float(np.random.rand()) * 0.3
        return sum(self.base[t] for t in rule) + noise

    def __getattr__(self, name):
        if name == "compute_relevances" and self.batch_api:
            def batched(pred, rules, snapshots=False):
                rels, snaps = [], []
                for r in rules:
                    rels.append(self.compute_relevance(pred, r))
                    snaps.append([torch.get_rng_state(), np.random.get_state()])
                return (rels, snaps) if snapshots else rels
            return batched
        raise AttributeError(name)


def sequential_reference(engine, xsi, pred, triples, length_cap=4, window=10, k=10):
    """Literal restatement of the reference control flow, one candidate at a time."""
    t2r = {t: engine.compute_relevance(pred, [t]) for t in triples}
    rule_to_rel = [((t,), r) for t, r in sorted(t2r.items(), key=lambda x: x[1], reverse=True)]
    n_rel, best = len(t2r), rule_to_rel[0][1]
    if not best > xsi:
        for length in range(2, min(len(t2r), length_cap) + 1):
            rules = sorted(((r, sum(t2r[t] for t in r)) for r in itertools.combinations(triples, length)),
                           key=lambda x: x[1], reverse=True)
            terminate, cbest, win, cur = False, -1e6, [None] * window, {}
            for i, (rule, _) in enumerate(rules):
                if terminate:
                    break
                rel = engine.compute_relevance(pred, list(rule))
                cur[rule] = rel
                n_rel += 1
                win[i % window] = rel
                if rel > xsi:
                    break
                elif rel >= cbest:
                    cbest = rel
                elif i >= window:
                    terminate = random.random() > (sum(win) / window) / cbest
            cur = sorted(cur.items(), key=lambda x: x[1], reverse=True)
            rule_to_rel += cur
            best = max(best, cur[0][1])
            if best > xsi:
                break
    rule_to_rel = sorted(rule_to_rel, key=lambda x: (x[1], 1 / len(x[0])), reverse=True)[:k]
    return rule_to_rel, n_rel


@pytest.mark.parametrize("batch_api", [True, False])
@pytest.mark.parametrize("xsi,seed", [(5.0, 1), (2.2, 2), (100.0, 3), (1.4, 4)])
def test_batched_builder_equals_sequential_loop(xsi, seed, batch_api):
    triples = [(1, 0, i) for i in range(9)]
    rng = np.random.default_rng(seed)
    base = {t: float(rng.random()) for t in triples}
    pred = (1, 0, 99)

    def seeds():
        torch.manual_seed(seed)
        np.random.seed(seed)
        random.seed(seed)

    seeds()
    ref_engine = RngEngine(base, False)
    ref_rules, ref_n = sequential_reference(ref_engine, xsi, pred, triples)
    ref_state = (torch.rand(1).item(), np.random.rand(), random.random())

    seeds()
    eng = RngEngine(base, batch_api)
    out = StochasticBuilder(xsi, eng, batch_size=7).build_explanations(pred, triples)
    state = (torch.rand(1).item(), np.random.rand(), random.random())

    assert out["#relevances"] == ref_n
    got = [(tuple(r), rel) for r, rel in out["rule_to_relevance"]]
    assert got == [(tuple(r), rel) for r, rel in ref_rules]
    assert state == ref_state  # generators end exactly where the sequential loop leaves them


@pytest.mark.parametrize("xsi,seed", [(5.0, 11), (2.0, 12)])
def test_batched_builder_equals_the_reference_class(xsi, seed, capsys):
    """Same check against the UNMODIFIED reference StochasticBuilder (build container only)."""
    import os
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference not mounted")
    from oracle import refshim
    refshim.install(cpu=False)
    from src.explanation_builders.stochastic_builder import StochasticBuilder as RefBuilder

    class DS(FakeDataset):
        def printable_nple(self, nple):
            return ""

    triples = [(1, 0, i) for i in range(8)]
    rng = np.random.default_rng(seed)
    base = {t: float(rng.random()) for t in triples}
    pred = (1, 0, 99)
    outs = []
    for cls, kw in ((RefBuilder, {}), (StochasticBuilder, {"batch_size": 5})):
        torch.manual_seed(seed); np.random.seed(seed); random.seed(seed)
        eng = RngEngine(base, cls is StochasticBuilder)
        eng.dataset = DS()
        out = cls(xsi, eng, **kw).build_explanations(pred, list(triples))
        outs.append((out["#relevances"], [(tuple(map(tuple, r)), rel) for r, rel in out["rule_to_relevance"]],
                     torch.rand(1).item(), random.random()))
    assert outs[0] == outs[1]
