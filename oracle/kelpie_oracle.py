"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's relevance-engine hot path.

This file is the *checker*: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The product
(`kelpie_b200/`) never does, and fails loudly when its CUDA library is missing.

It restates, with plain torch-CPU ops + autograd (the reference's arithmetic IS torch,
SURVEY.md section 8c), the algorithm of rbarile17/kelpie for:

  * mimic post-training         TransE  pairwise_ranking_optimizer.py:55-98,139-203
                                ComplEx multiclass_nll_optimizer.py:57-99,123-164
                                ConvE   bce_optimizer.py:45-112,161-208
  * the three score functions   transe.py:48-75, complex.py:59-113, conve.py:133-158
  * regularisers                regularizers.py:15-46
  * mimic construction          transe.py:84-99, complex.py:144-160, conve.py:193-237
  * mimic dataset overlay       kelpie_dataset.py:10-203
  * filtered rank (engine)      post_training_engine.py:101-125
  * relevance                   post_training_engine.py:46-62,132-191
  * predict_tails/_triples      model.py:25-68, conve.py:160-184
  * select_entities_to_convert  engine.py:22-126

Random numbers are consumed in EXACTLY the reference's order so
that, from the same seeds, this file reproduces the reference's outputs; that is how
it is pinned (tests/test_oracle_golden.py against tests/golden/*.npz, which
tests/golden/make_golden.py produced by running the unmodified reference here).
Parity is therefore pinned by outputs of the reference itself; the reference's own
test-suite holds no numeric golden vectors for this path.
"""
import math
import random
from collections import defaultdict

import numpy as np
import torch
import torch.nn.functional as F

ONE_TO_ONE, ONE_TO_MANY, MANY_TO_ONE, MANY_TO_MANY = "1-1", "1-N", "N-1", "N-N"


# --------------------------------------------------------------------------- data


class KG:
    """Id-triple graph plus the filter multisets of dataset.py:101-139."""

    def __init__(self, train, valid, test, num_entities, num_relations):
        self.train = np.asarray(train, dtype=np.int64).reshape(-1, 3)
        self.valid = np.asarray(valid, dtype=np.int64).reshape(-1, 3)
        self.test = np.asarray(test, dtype=np.int64).reshape(-1, 3)
        self.num_entities = int(num_entities)
        self.num_relations = int(num_relations)
        R = self.num_relations

        per_entity = defaultdict(list)
        for s, p, o in self.train:
            per_entity[s].append((s, p, o))
            per_entity[o].append((s, p, o))
        # dataset.py:113-116 -- de-duplicated through a Python set (order = set order)
        self.facts_of = defaultdict(list)
        for e, lst in per_entity.items():
            self.facts_of[e] = list(set(lst))
        self.degree = {e: len(v) for e, v in self.facts_of.items()}

        self.train_to_filter = defaultdict(list)
        for s, p, o in self.train:
            self.train_to_filter[(s, p)].append(o)
            self.train_to_filter[(o, p + R)].append(s)
        self.to_filter = defaultdict(list)
        for s, p, o in np.vstack([self.train, self.valid, self.test]):
            self.to_filter[(s, p)].append(o)
            self.to_filter[(o, p + R)].append(s)

        self.valid_of = defaultdict(list)
        self.test_of = defaultdict(list)
        for s, p, o in self.valid:
            self.valid_of[s].append((s, p, o))
            self.valid_of[o].append((s, p, o))
        for s, p, o in self.test:
            self.test_of[s].append((s, p, o))
            self.test_of[o].append((s, p, o))
        self._relation_types()

    def invert(self, triples):
        """dataset.py:319-331."""
        t = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
        out = t.copy()
        out[:, 0] = t[:, 2]
        out[:, 2] = t[:, 0]
        out[:, 1] = t[:, 1] + self.num_relations
        return out

    def _relation_types(self):
        """dataset.py:282-317."""
        R = self.num_relations
        s_num, o_num = defaultdict(list), defaultdict(list)
        for (e, r) in self.train_to_filter:
            n = len(self.to_filter[(e, r)])
            if r >= R:
                s_num[r - R].append(n)
            else:
                o_num[r].append(n)
        self.relation_to_type = {}
        for r in s_num:
            s_per_o, o_per_s = np.average(s_num[r]), np.average(o_num[r])
            if s_per_o > 1.2 and o_per_s > 1.2:
                self.relation_to_type[r] = MANY_TO_MANY
            elif s_per_o > 1.2:
                self.relation_to_type[r] = MANY_TO_ONE
            elif o_per_s > 1.2:
                self.relation_to_type[r] = ONE_TO_MANY
            else:
                self.relation_to_type[r] = ONE_TO_ONE


def _swap(triple, old, new):
    s, p, o = triple
    return (new if s == old else s, p, new if o == old else o)


class Mimic:
    """The overlay kelpie_dataset.py builds for one explained entity.

    Mimic id M = N (kelpie_dataset.py:20-25); facts = the entity's training facts with the
    entity renamed to M (:30-34); the filter multiset for keys that mention M
    (:52-62); removal / addition with one-occurrence semantics (:111-158).
    `facts` may be passed explicitly to fix the order (it is a Python-set order in the
    reference, dataset.py:113-116).
    """

    def __init__(self, kg, entity, facts=None):
        self.kg, self.entity, self.M = kg, int(entity), kg.num_entities
        R = kg.num_relations
        base = kg.facts_of[self.entity] if facts is None else facts
        self.base_facts = [_swap(t, self.entity, self.M) for t in base]
        self.facts = list(self.base_facts)
        extra = [_swap(t, self.entity, self.M) for t in kg.valid_of[self.entity]]
        extra += [_swap(t, self.entity, self.M) for t in kg.test_of[self.entity]]
        self.filter = defaultdict(list)
        for s, p, o in self.base_facts + extra:
            self.filter[(s, p)].append(o)
            self.filter[(o, p + R)].append(s)

    def as_mimic(self, triple):
        return _swap(triple, self.entity, self.M)

    def filter_for(self, s, p):
        if s == self.M:
            return list(self.filter.get((s, p), []))
        return list(self.kg.to_filter.get((s, p), [])) + list(self.filter.get((s, p), []))

    def without(self, rule):
        """remove_training_triples (kelpie_dataset.py:130-158): new fact list + filter."""
        R = self.kg.num_relations
        gone = [_swap(t, self.entity, self.M) for t in rule]
        idx = [self.base_facts.index(t) for t in gone]
        facts = [t for i, t in enumerate(self.base_facts) if i not in set(idx)]
        flt = {k: list(v) for k, v in self.filter.items()}
        for s, p, o in gone:
            flt[(s, p)].remove(o)
            flt[(o, p + R)].remove(s)
        return facts, flt

    def with_added(self, rule):
        """add_training_triples (kelpie_dataset.py:98-128)."""
        R = self.kg.num_relations
        new = [_swap(t, self.entity, self.M) for t in rule]
        facts = list(self.base_facts) + new
        flt = defaultdict(list, {k: list(v) for k, v in self.filter.items()})
        for s, p, o in new:
            flt[(s, p)].append(o)
            flt[(o, p + R)].append(s)
        return facts, flt


# --------------------------------------------------------------------------- models


class Weights:
    """Frozen parameters of a trained link-prediction model."""

    def __init__(self, kind, ent, rel, **kw):
        self.kind = kind  # "TransE" | "ComplEx" | "ConvE"
        self.ent = torch.as_tensor(ent, dtype=torch.float32)
        self.rel = torch.as_tensor(rel, dtype=torch.float32)
        self.norm = int(kw.get("norm", 2))
        self.init_scale = float(kw.get("init_scale", 1e-3))
        self.conve = kw.get("conve")  # dict of tensors, see conve_features
        self.dropout = kw.get("dropout", (0.0, 0.0, 0.0))  # input, feature-map, hidden

    @property
    def is_minimizer(self):
        return self.kind == "TransE"  # transe.py:35-36, complex.py:38-39, conve.py:62-63

    @property
    def dim(self):
        return self.ent.shape[1]


def conve_features(w, lhs, rel, training=False):
    """conve.py:133-156 up to (excluding) the entity projection.  BN in eval mode."""
    c = w.conve
    D = lhs.shape[1]
    width, height = 20, D // 20
    x = torch.cat([lhs.view(-1, 1, width, height), rel.view(-1, 1, width, height)], 2)
    x = F.batch_norm(x, c["bn1_mean"], c["bn1_var"], c["bn1_w"], c["bn1_b"], False, 0.1, 1e-5)
    x = F.dropout(x, w.dropout[0], training)
    x = F.conv2d(x, c["conv_w"], c["conv_b"])
    x = F.batch_norm(x, c["bn2_mean"], c["bn2_var"], c["bn2_w"], c["bn2_b"], False, 0.1, 1e-5)
    x = torch.relu(x)
    x = F.dropout2d(x, w.dropout[1], training)
    x = x.view(x.shape[0], -1)
    x = F.linear(x, c["fc_w"], c["fc_b"])
    x = F.dropout(x, w.dropout[2], training)
    x = F.batch_norm(x, c["bn3_mean"], c["bn3_var"], c["bn3_w"], c["bn3_b"], False, 0.1, 1e-5)
    return torch.relu(x)


def all_scores(w, table, triples, training=False):
    """Score (s,p,.) against every row of `table` -> [Q, rows]."""
    t = torch.as_tensor(np.asarray(triples), dtype=torch.long).view(-1, 3)
    lhs, rel = table[t[:, 0]], w.rel[t[:, 1]]
    if w.kind == "TransE":  # transe.py:48-65 (chunks of 2048 rows, norm over dim 2)
        out = []
        for chunk in torch.split(table, 2048, dim=0):
            diff = (lhs + rel).unsqueeze(0) - chunk.unsqueeze(1)
            out.append(diff.norm(p=w.norm, dim=2))
        return torch.cat(out, 0).transpose(0, 1)
    if w.kind == "ComplEx":  # complex.py:88-113 (chunks of 512 rows)
        d = w.dim // 2
        re = lhs[:, :d] * rel[:, :d] - lhs[:, d:] * rel[:, d:]
        im = lhs[:, :d] * rel[:, d:] + lhs[:, d:] * rel[:, :d]
        q = torch.cat([re, im], 1)
        return torch.cat([q @ c.transpose(0, 1) for c in torch.split(table, 512, dim=0)], 1)
    if w.kind == "ConvE":  # conve.py:133-158
        x = conve_features(w, lhs, rel, training)
        return torch.sigmoid(torch.mm(x, table.transpose(1, 0)))
    raise ValueError(w.kind)


def _l2(factors, weight):
    """regularizers.py:15-22."""
    return sum(torch.mean(f ** 2) for f in factors) * weight / len(factors)


def _n3(factors, weight):
    """regularizers.py:37-46."""
    return sum(weight * torch.sum(torch.abs(f) ** 3) for f in factors) / factors[0].shape[0]


def _n2(factors, weight):
    """regularizers.py:25-34."""
    return sum(weight * torch.sum(torch.norm(f, 2, 1) ** 3) for f in factors) / factors[0].shape[0]


def init_mimic_row(w, init_tensor):
    """Mimic-row initialisation of the three Kelpie model classes.

    TransE re-draws with xavier_normal_ (transe.py:92-95; consumes the generator of the
    tensor's device -- the CPU one here); ComplEx scales by init_scale
    (complex.py:155-157); ConvE uses the tensor as is (conve.py:209) -- but building the
    inner ConvE (conve.py:202 -> :50-52) constructs a throw-away nn.Conv2d and nn.Linear
    whose default initialisers DRAW FROM THE CPU GENERATOR before being replaced by
    deep copies (:214-222); those draws are reproduced here or every later random
    number of the run would differ from the reference's.
    """
    row = init_tensor.clone()
    if w.kind == "TransE":
        torch.nn.init.xavier_normal_(row)
    elif w.kind == "ComplEx":
        row *= w.init_scale
    elif w.kind == "ConvE":
        hidden, dim = w.conve["fc_w"].shape[1], w.conve["fc_w"].shape[0]
        torch.nn.Conv2d(1, w.conve["conv_w"].shape[0], (3, 3), 1, 0, bias=True)
        torch.nn.Linear(hidden, dim)
    return row


# --------------------------------------------------------------------------- post-training


def post_train(w, kg, mimic_row, facts, hp, log=None):
    """Embedding-only training of the mimic row over `facts`; returns the [N+1, D] table.

    The table is `cat([frozen entities, mimic row])` built once, with the mimic row
    written back in place after every optimiser step (model.py:110-112) -- the same
    autograd structure as the reference (dense [N+1, D] gradient through CatBackward).
    `log`, when a list, receives one dict per optimiser step describing the rows used
    (the pre-drawn inputs the CUDA path consumes).
    """
    param = torch.nn.Parameter(mimic_row.clone(), requires_grad=True)
    table = torch.cat([w.ent.clone().detach(), param], 0)
    M = w.ent.shape[0]
    n_ent = M + 1
    facts = np.asarray(facts, dtype=np.int64).reshape(-1, 3)
    if len(facts) == 0:
        rows = np.zeros((0, 3), dtype=np.int64)
    else:
        rows = np.vstack((facts, kg.invert(facts)))

    def write_back():
        with torch.no_grad():
            table[M] = param

    if w.kind == "TransE":
        # pairwise_ranking_optimizer.py:28-53 (Adam, MarginRankingLoss, L2)
        opt = torch.optim.Adam([param], lr=hp["lr"])
        margin, ratio, bs = hp["margin"], hp["negative_triples_ratio"], hp["batch_size"]
        for _ in range(hp["epochs"]):
            # :166-181 -- shuffle in place, repeat, randint(entities) THEN randint(2)
            np.random.shuffle(rows)
            pos = torch.from_numpy(np.repeat(rows, ratio, axis=0))
            size = (len(pos),)
            rnd = torch.randint(high=n_ent, size=size)
            coin = torch.randint(high=2, size=size)
            head = coin == 1
            neg = torch.stack(
                (torch.where(head, rnd, pos[:, 0]), pos[:, 1], torch.where(~head, rnd, pos[:, 2])), 1
            )
            # :187-201 -- only the first len(rows) of the ratio*len(rows) rows are visited
            for b0 in range(0, len(rows), bs):
                b1 = min(b0 + bs, len(rows))
                pb, nb = pos[b0:b1], neg[b0:b1]
                if log is not None:
                    log.append({"pos": pb.numpy().copy(), "neg": nb.numpy().copy()})
                opt.zero_grad()

                def fwd(t):  # transe.py:67-75
                    l, r, o = table[t[:, 0]], w.rel[t[:, 1]], table[t[:, 2]]
                    return (l + r - o).norm(p=w.norm, dim=1), (l, r, o)

                ps, pf = fwd(pb)
                ns, nf = fwd(nb)
                fit = F.margin_ranking_loss(ps, ns, torch.tensor([-1.0]), margin=margin)
                reg = (_l2(pf, hp["regularizer_weight"]) + _l2(nf, hp["regularizer_weight"])) / 2
                (fit + reg).backward()
                opt.step()
                write_back()
    elif w.kind == "ComplEx":
        # multiclass_nll_optimizer.py:27-55 (Adagrad by default, N3)
        name = hp.get("optimizer_name", "Adagrad")
        if name == "Adagrad":
            opt = torch.optim.Adagrad([param], lr=hp["lr"])
        elif name == "Adam":
            opt = torch.optim.Adam([param], lr=hp["lr"], betas=(hp["decay1"], hp["decay2"]))
        else:
            opt = torch.optim.SGD([param], lr=hp["lr"])
        rows_t = torch.from_numpy(rows)
        bs = min(hp["batch_size"], len(rows_t))
        d = w.dim // 2
        for _ in range(hp["epochs"]):
            perm = rows_t[torch.randperm(rows_t.shape[0]), :]  # :148
            b0 = 0
            while b0 < rows_t.shape[0]:  # :155-162
                batch = perm[b0 : b0 + bs]
                if log is not None:
                    log.append({"rows": batch.numpy().copy()})
                lhs, rel, rhs = table[batch[:, 0]], w.rel[batch[:, 1]], table[batch[:, 2]]
                lr_, li_ = lhs[:, :d], lhs[:, d:]
                rr_, ri_ = rel[:, :d], rel[:, d:]
                hr_, hi_ = rhs[:, :d], rhs[:, d:]
                # complex.py:59-86
                logits = (lr_ * rr_ - li_ * ri_) @ table[:, :d].transpose(0, 1) + (
                    lr_ * ri_ + li_ * rr_
                ) @ table[:, d:].transpose(0, 1)
                factors = (
                    torch.sqrt(lr_ ** 2 + li_ ** 2),
                    torch.sqrt(rr_ ** 2 + ri_ ** 2),
                    torch.sqrt(hr_ ** 2 + hi_ ** 2),
                )
                reg = {"N3": _n3, "N2": _n2}[hp.get("regularizer_name", "N3")]  # multiclass_nll_optimizer.py:46-49
                loss = F.cross_entropy(logits, batch[:, 2]) + reg(factors, hp["regularizer_weight"])
                opt.zero_grad()
                loss.backward()
                opt.step()
                write_back()
                b0 += hp["batch_size"]
    elif w.kind == "ConvE":
        # bce_optimizer.py:161-165 -- Adam re-created WITHOUT lr => lr = 1e-3 (config lr unused)
        opt = torch.optim.Adam([param])
        vocab = defaultdict(list)  # :92-96, first-seen key order
        for s, p, o in rows:
            vocab[(int(s), int(p))].append(int(o))
        pairs = list(vocab.keys())
        bs, ls = hp["batch_size"], hp["label_smoothing"]
        for _ in range(hp["epochs"]):
            for b0 in range(0, len(pairs), bs):
                batch = pairs[b0 : b0 + bs]
                targets = torch.zeros((len(batch), n_ent))  # :98-112
                for i, pair in enumerate(batch):
                    targets[i, vocab[pair]] = 1.0
                if ls:
                    targets = (1.0 - ls) * targets
                    targets += 1.0 / targets.shape[1]
                if log is not None:
                    log.append({"pairs": np.array(batch, dtype=np.int64)})
                opt.zero_grad()
                bt = np.array([(s, p, 0) for s, p in batch], dtype=np.int64)
                pred = all_scores(w, table, bt, training=True)
                F.binary_cross_entropy(pred, targets).backward()
                opt.step()
                write_back()
    else:
        raise ValueError(w.kind)
    return table.detach()


# --------------------------------------------------------------------------- ranks


def triple_results(w, table, triple, filter_out):
    """post_training_engine.py:101-125 (asymmetric min/max semantics kept)."""
    s, p, o = triple
    with torch.no_grad():
        scores = all_scores(w, table, np.array([triple]))[0].detach().clone()
    target = scores[o].item()
    idx = torch.as_tensor(list(filter_out), dtype=torch.long)
    if w.is_minimizer:
        scores[idx] = 1e6
        scores[o] = target
        best = torch.min(scores)
        rank = torch.sum(scores <= target)
    else:
        scores[idx] = -1e6
        best = torch.max(scores)
        rank = torch.sum(scores >= target)
    return {"target_score": target, "best_score": float(best), "target_rank": int(rank)}


def predict_tails(w, kg, triples):
    """model.py:42-68; ConvE override conve.py:160-184 (filter value 0.0, sort position)."""
    triples = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
    scores_out, ranks = [], []
    with torch.no_grad():
        if w.kind == "ConvE":
            for i in range(0, len(triples), 128):
                batch = triples[i : i + 128]
                sc = all_scores(w, w.ent, batch).clone()
                for j, (s, p, o) in enumerate(batch):
                    flt = kg.to_filter[(s, p)]
                    t = sc[j, o].item()
                    scores_out.append(t)
                    sc[j, flt] = 0.0
                    sc[j, o] = t
                order = torch.sort(sc, dim=1, descending=True)[1].numpy()
                for j in range(len(batch)):
                    ranks.append(int(np.where(order[j] == batch[j, 2])[0][0]) + 1)
            return scores_out, ranks
        sc = all_scores(w, w.ent, triples).clone()
        targets = torch.zeros((len(triples), 1))
        for i, (_, _, o) in enumerate(triples):
            targets[i, 0] = sc[i, o].item()
        default = 1e6 if w.is_minimizer else -1e6
        for i, (s, p, o) in enumerate(triples):
            flt = list(set(kg.to_filter[(s, p)]))
            sc[i, torch.as_tensor(flt, dtype=torch.long)] = default
            sc[i, o] = targets[i, 0]
        cmp = (sc <= targets) if w.is_minimizer else (sc >= targets)
        ranks = torch.sum(cmp.float(), dim=1).numpy().tolist()
        scores_out = [float(targets[i, 0]) for i in range(len(triples))]
    return scores_out, ranks


def predict_triples(w, kg, triples):
    """model.py:25-40."""
    triples = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
    ds, tr = predict_tails(w, kg, triples)
    hs, hr = predict_tails(w, kg, kg.invert(triples))
    return [
        {"score": {"tail": ds[i], "head": hs[i]}, "rank": {"tail": int(tr[i]), "head": int(hr[i])}}
        for i in range(len(triples))
    ]


def convertible_entities(w, kg, pred, degree_cap=None):
    """engine.py:62-124 -- every eligible head whose (e,p,o) is not already rank 1."""
    s, p, o = pred
    eligible = []
    for e in range(kg.num_entities):
        if e == s or kg.degree.get(e, 0) < 1:
            continue
        if degree_cap and kg.degree[e] > degree_cap:
            continue
        if (e, p) in kg.to_filter:
            if kg.relation_to_type[p] in (ONE_TO_ONE, MANY_TO_ONE):
                continue
            if o in kg.to_filter[(e, p)]:
                continue
        eligible.append(e)
    out = []
    for b0 in range(0, len(eligible), 4):
        chunk = eligible[b0 : b0 + 4]
        with torch.no_grad():
            sc = all_scores(w, w.ent, np.array([(e, p, o) for e in chunk])).numpy().copy()
        for j, e in enumerate(chunk):
            row = sc[j]
            flt = np.array(kg.to_filter.get((e, p), []), dtype=np.int64)
            t = row[o]
            if w.is_minimizer:
                row[flt] = 1e6
                if 1e6 > t > np.min(row):
                    out.append(e)
            else:
                row[flt] = -1e6
                if -1e6 < t < np.max(row):
                    out.append(e)
    return out


def select_entities_to_convert(w, kg, pred, k, degree_cap=None):
    """engine.py:125 -- random.sample on the Python generator."""
    pool = convertible_entities(w, kg, pred, degree_cap)
    return random.sample(pool, k=min(k, len(pool)))


# --------------------------------------------------------------------------- engines


def _sigmoid(x):
    return 1 / (1 + math.exp(-x))  # post_training_engine.py:18-20


class Engine:
    """Necessary / sufficient post-training engines (post_training_engine.py:17-207)."""

    def __init__(self, w, kg, hp, mode="necessary", fact_order=None):
        self.w, self.kg, self.hp, self.mode = w, kg, dict(hp), mode
        self.fact_order = fact_order or {}
        self.entities_to_convert = []
        self.set_cache()

    def set_cache(self):
        self.base = {}
        self.mimics = {}
        self.trace = []  # (tag, mimic_init, mimic_final, results) per post-training

    def _mimic(self, e):
        if e not in self.mimics:
            self.mimics[e] = Mimic(self.kg, e, self.fact_order.get(e))
        return self.mimics[e]

    def _results(self, tag, row, facts, flt, mimic, pred, log=None):
        table = post_train(self.w, self.kg, row, facts, self.hp, log)
        mp = mimic.as_mimic(pred)
        res = triple_results(self.w, table, mp, flt.get((mp[0], mp[1]), []))
        self.trace.append((tag, row.clone(), table[-1].clone(), res))
        return res

    def individual(self, pred, rule, log=None):
        """post_training_engine.py:46-62: returns (pt_results, base_results)."""
        pred = tuple(int(x) for x in pred)
        mimic = self._mimic(pred[0])
        init = torch.rand(1, self.w.dim)  # :52, CPU generator
        base_row = init_mimic_row(self.w, init)  # :55 -- always built, even if cached
        if pred not in self.base:  # :83-88
            self.base[pred] = self._results("base", base_row, mimic.base_facts, mimic.filter, mimic, pred)
        pt_row = init_mimic_row(self.w, init)  # :59
        rule = [tuple(int(x) for x in t) for t in rule]
        facts, flt = mimic.without(rule) if self.mode == "necessary" else mimic.with_added(rule)
        pt = self._results("pt", pt_row, facts, flt, mimic, pred, log)
        return pt, self.base[pred]

    def compute_relevance(self, pred, rule):
        if self.mode == "necessary":  # :132-145
            pt, base = self.individual(pred, rule)
            d_rank = torch.tensor(pt["target_rank"]) - torch.tensor(base["target_rank"])
            d = pt["target_score"] - base["target_score"]
            if not self.w.is_minimizer:
                d = -d
            return float(d_rank + _sigmoid(d))
        rels = []  # :178-191
        s = pred[0]
        for e in self.entities_to_convert:
            c_rule = [_swap(t, s, e) for t in rule]
            c_pred = _swap(pred, s, e)
            pt, base = self.individual(c_pred, c_rule)
            d_rank = torch.tensor(base["target_rank"]) - torch.tensor(pt["target_rank"])
            d = base["target_score"] - pt["target_score"]
            if not self.w.is_minimizer:
                d = -d
            rel = float(d_rank + _sigmoid(d))
            rels.append(rel / float(base["target_rank"]))
        return sum(rels) / len(rels)


def train_transe_full(ent, rel, norm, training_triples, num_entities, num_relations, hp, n_epochs=None):
    """Full-model TransE training: PairwiseRankingOptimizer.train / epoch / step_on_batch
    (pairwise_ranking_optimizer.py:55-157) restated with torch autograd on the CPU; consumes np.random and the
    torch CPU generator in the reference's order (shuffle, randint(2), randint(num_entities) per epoch).
    ent / rel: float32 arrays; returns the trained copies."""
    E = torch.nn.Parameter(torch.from_numpy(np.array(ent, dtype=np.float32)))
    R = torch.nn.Parameter(torch.from_numpy(np.array(rel, dtype=np.float32)))
    opt = torch.optim.Adam([E, R], lr=hp["lr"])
    loss_fn = torch.nn.MarginRankingLoss(margin=hp["margin"], reduction="mean")
    t = np.asarray(training_triples).astype(np.int64).reshape(-1, 3)
    inv = t.copy()
    inv[:, 0], inv[:, 2] = t[:, 2], t[:, 0]
    inv[:, 1] = t[:, 1] + num_relations
    rows = np.vstack((t, inv))
    ratio, bs, w = int(hp["negative_triples_ratio"]), int(hp["batch_size"]), float(hp["regularizer_weight"])

    def fwd(b):
        lhs, rl, rhs = E[b[:, 0]], R[b[:, 1]], E[b[:, 2]]
        return (lhs + rl - rhs).norm(p=norm, dim=1), (lhs, rl, rhs)

    for _ in range(int(n_epochs if n_epochs is not None else hp["epochs"])):
        np.random.shuffle(rows)
        pos = torch.from_numpy(np.repeat(rows, ratio, axis=0))
        size = torch.Size([len(pos)])
        coin = torch.randint(high=2, size=size)
        rnd = torch.randint(high=num_entities, size=size)
        head = coin == 1
        neg = torch.stack((torch.where(head, rnd, pos[:, 0]), pos[:, 1], torch.where(~head, rnd, pos[:, 2])), dim=1)
        for b0 in range(0, len(rows), bs):
            b1 = min(b0 + bs, len(rows))
            opt.zero_grad()
            ps, pf = fwd(pos[b0:b1])
            ns, nf = fwd(neg[b0:b1])
            loss = loss_fn(ps, ns, torch.tensor([-1.0])) + (_l2(pf, w) + _l2(nf, w)) / 2
            loss.backward()
            opt.step()
    return E.detach().numpy(), R.detach().numpy()


def train_complex_full(ent, rel, training_triples, num_relations, hp, n_epochs=None):
    """Full-model ComplEx training: MultiClassNLLOptimizer.train / epoch / step_on_batch
    (multiclass_nll_optimizer.py:58-135) with ComplEx.forward (complex.py:58-86), restated with torch autograd on
    the CPU; consumes the torch CPU generator like the reference (one randperm per epoch)."""
    E = torch.nn.Parameter(torch.from_numpy(np.array(ent, dtype=np.float32)))
    R = torch.nn.Parameter(torch.from_numpy(np.array(rel, dtype=np.float32)))
    name = hp["optimizer_name"]
    if name == "Adam":
        opt = torch.optim.Adam([E, R], lr=hp["lr"], betas=(hp["decay1"], hp["decay2"]))
    else:
        opt = {"Adagrad": torch.optim.Adagrad, "SGD": torch.optim.SGD}[name]([E, R], lr=hp["lr"])
    reg = {"N3": _n3, "N2": _n2}[hp["regularizer_name"]]
    t = np.asarray(training_triples).astype(np.int64).reshape(-1, 3)
    inv = t.copy()
    inv[:, 0], inv[:, 2] = t[:, 2], t[:, 0]
    inv[:, 1] = t[:, 1] + num_relations
    rows = torch.from_numpy(np.vstack((t, inv)))
    d = E.shape[1] // 2
    bs = min(int(hp["batch_size"]), len(rows))
    loss_fn = torch.nn.CrossEntropyLoss(reduction="mean")
    for _ in range(int(n_epochs if n_epochs is not None else hp["epochs"])):
        perm = rows[torch.randperm(rows.shape[0]), :]
        b0 = 0
        while b0 < len(perm):
            b = perm[b0:min(b0 + bs, len(perm))]
            lhs, rl, rhs = E[b[:, 0]], R[b[:, 1]], E[b[:, 2]]
            lhs, rl, rhs = (lhs[:, :d], lhs[:, d:]), (rl[:, :d], rl[:, d:]), (rhs[:, :d], rhs[:, d:])
            score = (lhs[0] * rl[0] - lhs[1] * rl[1]) @ E[:, :d].t() + (lhs[0] * rl[1] + lhs[1] * rl[0]) @ E[:, d:].t()
            factors = (torch.sqrt(lhs[0] ** 2 + lhs[1] ** 2), torch.sqrt(rl[0] ** 2 + rl[1] ** 2), torch.sqrt(rhs[0] ** 2 + rhs[1] ** 2))
            loss = loss_fn(score, b[:, 2]) + reg(factors, float(hp["regularizer_weight"]))
            opt.zero_grad()
            loss.backward()
            opt.step()
            b0 += int(hp["batch_size"])
    return E.detach().numpy(), R.detach().numpy()


CONVE_STATE_KEYS = ("entity_embeddings", "relation_embeddings", "batch_norm_1.weight", "batch_norm_1.bias",
                    "batch_norm_1.running_mean", "batch_norm_1.running_var", "batch_norm_2.weight", "batch_norm_2.bias",
                    "batch_norm_2.running_mean", "batch_norm_2.running_var", "batch_norm_3.weight", "batch_norm_3.bias",
                    "batch_norm_3.running_mean", "batch_norm_3.running_var", "convolutional_layer.weight",
                    "convolutional_layer.bias", "hidden_layer.weight", "hidden_layer.bias")


def train_conve_full(state, training_triples, num_entities, num_relations, hp, n_epochs=None, max_steps=None, dtype=np.float32):
    """Full-model ConvE training: BCEOptimizer.train / epoch / extract_batch / step_on_batch (bce_optimizer.py:44-158)
    with ConvE.forward = all_scores (conve.py:133-158), restated with torch autograd on the CPU (dropout rates 0).
    `state`: dict of arrays under the reference's state-dict keys (CONVE_STATE_KEYS); returns the trained dict.
    Quirks kept: pairs (s, p) in first-appearance order over triples + inverses, np.random.shuffle of the PAIR LIST
    per epoch (cumulative), multi-hot targets * (1 - ls) + 1 / N, batch-norm in train mode (batch statistics, running
    statistics with momentum 0.1 and the unbiased variance) except on a step of ONE pair (:140-156), Adam over every
    parameter, ExponentialLR stepped per epoch when decay != 0 (:125-126)."""
    P = {k: torch.nn.Parameter(torch.from_numpy(np.array(state[k], dtype=dtype))) for k in CONVE_STATE_KEYS
         if "running" not in k}
    S = {k: torch.from_numpy(np.array(state[k], dtype=dtype)) for k in CONVE_STATE_KEYS if "running" in k}
    opt = torch.optim.Adam(list(P.values()), lr=hp["lr"])
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, hp["decay"])
    t = np.asarray(training_triples).astype(np.int64).reshape(-1, 3)
    inv = t.copy()
    inv[:, 0], inv[:, 2] = t[:, 2], t[:, 0]
    inv[:, 1] = t[:, 1] + num_relations
    vocab = {}
    for s, p, o in np.vstack((t, inv)):
        vocab.setdefault((s, p), []).append(o)
    pairs = list(vocab.keys())
    bs, ls, N = int(hp["batch_size"]), float(hp["label_smoothing"]), int(num_entities)
    D = P["entity_embeddings"].shape[1]
    loss_fn = torch.nn.BCELoss()

    def bn(x, name, train):
        return F.batch_norm(x, S[name + ".running_mean"], S[name + ".running_var"], P[name + ".weight"], P[name + ".bias"],
                            train, 0.1, 1e-5)

    done = 0
    for _ in range(int(n_epochs if n_epochs is not None else hp["epochs"])):
        np.random.shuffle(pairs)
        for b0 in range(0, len(pairs), bs):
            if max_steps is not None and done >= max_steps:  # diagnostic: stop after a number of steps
                break
            done += 1
            batch = pairs[b0:b0 + bs]
            targets = torch.zeros((len(batch), N), dtype=P["entity_embeddings"].dtype)
            for i, pr in enumerate(batch):
                targets[i, vocab[pr]] = 1.0
            if ls:
                targets = (1.0 - ls) * targets + 1.0 / N
            idx = torch.tensor(batch)
            train = len(batch) > 1
            lhs = P["entity_embeddings"][idx[:, 0]].view(-1, 1, 20, D // 20)
            rel = P["relation_embeddings"][idx[:, 1]].view(-1, 1, 20, D // 20)
            x = bn(torch.cat([lhs, rel], 2), "batch_norm_1", train)
            x = F.conv2d(x, P["convolutional_layer.weight"], P["convolutional_layer.bias"])
            x = torch.relu(bn(x, "batch_norm_2", train)).view(len(batch), -1)
            x = F.linear(x, P["hidden_layer.weight"], P["hidden_layer.bias"])
            x = torch.relu(bn(x, "batch_norm_3", train))
            pred = torch.sigmoid(x @ P["entity_embeddings"].t())
            loss = loss_fn(pred, targets)
            opt.zero_grad()
            loss.backward()
            opt.step()
        if hp["decay"]:
            sched.step()
    out = {k: v.detach().numpy() for k, v in P.items()}
    out.update({k: v.numpy() for k, v in S.items()})
    return out


def dp_relevance(ent, rel, pred, fact, entity, epsilon, sufficient=False, lambd=1.0):
    """Data-poisoning relevance for ComplEx (data_poisoning_engine.py:21-50 get_gradient via autograd, :53-94 necessary,
    :97-131 sufficient), restated with torch autograd on the CPU in fp32 like the reference."""
    E = torch.as_tensor(ent, dtype=torch.float32)
    R = torch.as_tensor(rel, dtype=torch.float32)
    d = E.shape[1] // 2

    def score(lhs, r, rhs):  # complex.py:47-56
        lhs, r, rhs = (lhs[:, :d], lhs[:, d:]), (r[:, :d], r[:, d:]), (rhs[:, :d], rhs[:, d:])
        return torch.sum((lhs[0] * r[0] - lhs[1] * r[1]) * rhs[0] + (lhs[0] * r[1] + lhs[1] * r[0]) * rhs[1], 1, keepdim=True)

    ps, pp, po = pred
    lhs, r, rhs = E[ps].clone().view(1, -1), R[pp].view(1, -1), E[po].clone().view(1, -1)
    x = lhs if entity == ps else rhs
    x.requires_grad = True
    score(lhs, r, rhs).backward()
    g = x.grad[0]
    pert = E[entity] + epsilon * g if sufficient else E[entity] - epsilon * g  # ComplEx maximises
    s, p, o = fact
    L, Rr, O = E[[s, s]].clone(), R[[p, p]], E[[o, o]].clone()
    if s == entity:
        L[1] = pert
    else:
        O[1] = pert
    sc = score(L, Rr, O).detach().numpy()
    a, b = sc[0], sc[1]
    return float((-a + lambd * b)[0]) if sufficient else float((a - lambd * b)[0])
