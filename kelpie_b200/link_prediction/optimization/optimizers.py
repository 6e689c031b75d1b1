"""`Optimizer(model, hp, verbose).train(training_triples)` interface of the reference
(src/link_prediction/optimization/*.py).  Only the Kelpie* subclasses -- the mimic
post-training -- are on the hot path; they run as ONE job through the batched CUDA kernels
(the engines batch many).  Of the full-model trainers TransE's and ComplEx's run on the
device (PairwiseRankingOptimizer.train, MultiClassNLLOptimizer.train, BCEOptimizer.train).
"""
import numpy as np
import torch
from pydantic import BaseModel

from ... import plans, runtime
from ..models.model import KelpieModel, context_for


class PairwiseRankingOptimizerHyperParams(BaseModel):
    batch_size: int
    epochs: int
    lr: float
    margin: float
    negative_triples_ratio: int
    regularizer_weight: float


class MultiClassNLLOptimizerHyperParams(BaseModel):
    optimizer_name: str
    batch_size: int
    epochs: int
    lr: float
    decay1: float
    decay2: float
    regularizer_name: str
    regularizer_weight: float


class BCEOptimizerHyperParams(BaseModel):
    batch_size: int
    label_smoothing: float
    lr: float
    decay: float
    epochs: int


class Optimizer:
    def __init__(self, model, hp, verbose: bool = True):
        self.model = model
        self.dataset = self.model.dataset
        self.verbose = verbose
        self.hp = hp.model_dump() if isinstance(hp, BaseModel) else dict(hp)

    def train(self, training_triples, save_path=None, evaluate_every=-1, valid_triples=None):
        raise NotImplementedError(
            "training every embedding of a model is outside the accelerated path; use the reference's trainer")


class _KelpieOptimizer(Optimizer):
    """One mimic post-training = a batch of one job."""

    def train(self, training_triples, save_path=None, evaluate_every=-1, valid_triples=None):
        if not isinstance(self.model, KelpieModel):
            raise Exception("Kelpie optimizers post-train a KelpieModel")
        base = self.model.model
        ctx = context_for(base)
        ds = self.model.dataset  # KelpieDataset: num_entities = N + 1
        batch = plans.Batch(base.name, ds.num_entities - 1, ds.num_relations, self.hp)
        batch.add(np.array(training_triples).reshape(-1, 3), self.model.kelpie_entity_emb.detach().cpu().numpy())
        rows = ctx.post_train(runtime.make_hp(base.name, self.hp), **batch.arrays(compact=True))
        with torch.no_grad():
            self.model.kelpie_entity_emb = rows[:1].clone()
        self.model.update_embeddings()


class PairwiseRankingOptimizer(Optimizer):
    """Full-model TransE training (pairwise_ranking_optimizer.py:55-157), used by verify_explanations to
    retrain from scratch: every epoch's shuffle / corruptions are drawn on the host in the reference's
    order, the steps (margin-ranking loss + L2 + Adam over both tables) run in kp_transe_fit_steps."""

    def get_hyperparams_class():
        return PairwiseRankingOptimizerHyperParams

    def get_kelpie_class():
        return KelpiePairwiseRankingOptimizer

    def train(self, training_triples, save_path=None, eval_every=-1, valid_triples=None, trial=None, patience=5):
        hp, model = self.hp, self.model
        if isinstance(model, KelpieModel):
            raise Exception("the full-model trainer does not post-train a KelpieModel")
        if not torch.cuda.is_available():
            raise RuntimeError("kelpie_b200 has no CPU training path")
        rows = np.vstack((np.asarray(training_triples), self.dataset.invert_triples(training_triples))).astype(np.int64)
        model.invalidate_context()  # the scoring context borrows the tables that are about to change
        model.cuda()
        ent = model.entity_embeddings.data.contiguous()
        rel = model.relation_embeddings.data.contiguous()
        fit = runtime.TransEFit(ent, rel, model.norm, hp["lr"], hp["margin"], hp["regularizer_weight"])
        n, bs, ratio = len(rows), int(hp["batch_size"]), int(hp["negative_triples_ratio"])
        off = np.append(np.arange(0, n, bs), n).astype(np.int64)
        best, bad = None, 0
        self.epoch_losses = []
        try:
            for e in range(1, int(hp["epochs"]) + 1):
                pos, neg = plans.draw_transe_full_epoch(rows, self.dataset.num_entities, ratio)
                loss = fit.steps(pos, neg, off, want_loss=self.verbose)
                if loss is not None:
                    self.epoch_losses.append(float(loss.mean()))
                if valid_triples is not None and eval_every > 0 and e % eval_every == 0:
                    from ..evaluation import Evaluator
                    torch.cuda.synchronize()
                    h1 = Evaluator(model).evaluate(valid_triples)["h1"]
                    model.invalidate_context()
                    if trial is not None:
                        trial.report(h1, e)
                        if trial.should_prune():
                            raise RuntimeError("trial pruned")
                    if best is None or h1 > best:
                        best, bad = h1, 0
                    else:
                        bad += 1
                    if bad >= patience:
                        break
            self.launches = fit.launches()
        finally:
            fit.close()
        model.entity_embeddings.data, model.relation_embeddings.data = ent, rel
        model.invalidate_context()
        if save_path is not None:
            torch.save(model.state_dict(), save_path)


class KelpiePairwiseRankingOptimizer(_KelpieOptimizer):
    """pairwise_ranking_optimizer.py:160-203."""


class MultiClassNLLOptimizer(Optimizer):
    """Full-model ComplEx training (multiclass_nll_optimizer.py:58-135), used by verify_explanations to retrain
    from scratch: every epoch's permutation is drawn on the host (torch.randperm, the reference's order), the
    steps (1-vs-all cross-entropy against the whole entity table + Adagrad / Adam / SGD over both tables) run in
    kp_complex_fit_steps on the tensor cores."""

    def get_hyperparams_class():
        return MultiClassNLLOptimizerHyperParams

    def get_kelpie_class():
        return KelpieMultiClassNLLOptimizer

    def train(self, training_triples, save_path=None, eval_every=-1, valid_triples=None, trial=None, patience=5):
        hp, model = self.hp, self.model
        if isinstance(model, KelpieModel):
            raise Exception("the full-model trainer does not post-train a KelpieModel")
        if not torch.cuda.is_available():
            raise RuntimeError("kelpie_b200 has no CPU training path")
        rows = np.vstack((np.asarray(training_triples), self.dataset.invert_triples(training_triples))).astype(np.int64)
        n = len(rows)
        bs = min(int(hp["batch_size"]), n)  # :70
        model.invalidate_context()
        model.cuda()
        ent = model.entity_embeddings.data.contiguous()
        rel = model.relation_embeddings.data.contiguous()
        if hp.get("regularizer_name", "N3") != "N3" and float(hp["regularizer_weight"]) != 0:
            # multiclass_nll_optimizer.py:46-49 also offers N2; the full-model trainer's kernels apply N3 only (the
            # mimic post-training implements both) -- refuse rather than silently train another model
            raise NotImplementedError("full-model ComplEx training implements the N3 regulariser only")
        fit = runtime.ComplExFit(ent, rel, hp["optimizer_name"], hp["lr"], hp["decay1"], hp["decay2"],
                                 hp["regularizer_weight"], bs)
        starts = np.arange(0, n, int(hp["batch_size"]))  # batch_start += self.batch_size (:119)
        off = np.append(starts, n).astype(np.int64)
        keep = np.minimum(off[1:], off[:-1] + bs) == off[1:]
        assert keep.all()
        best, bad = None, 0
        self.epoch_losses = []
        try:
            for e in range(1, int(hp["epochs"]) + 1):
                perm = torch.randperm(n).numpy()
                loss = fit.steps(rows[perm].astype(np.int32), off, want_loss=self.verbose)
                if loss is not None:
                    self.epoch_losses.append(float(loss.mean()))
                if valid_triples is not None and eval_every > 0 and (e + 1) % eval_every == 0:  # :76
                    from ..evaluation import Evaluator
                    torch.cuda.synchronize()
                    h1 = Evaluator(model).evaluate(valid_triples)["h1"]
                    model.invalidate_context()
                    if trial is not None:
                        trial.report(h1, e)
                        if trial.should_prune():
                            raise RuntimeError("trial pruned")
                    if best is None or h1 > best:
                        best, bad = h1, 0
                    else:
                        bad += 1
                    if bad >= patience:
                        break
            self.launches = fit.launches()
        finally:
            fit.close()
        model.entity_embeddings.data, model.relation_embeddings.data = ent, rel
        model.invalidate_context()
        if save_path is not None:
            torch.save(model.state_dict(), save_path)


class KelpieMultiClassNLLOptimizer(_KelpieOptimizer):
    """multiclass_nll_optimizer.py:138-164."""


class BCEOptimizer(Optimizer):
    """Full-model ConvE training (bce_optimizer.py:44-158), used by verify_explanations to retrain from scratch: the
    (s, p) -> objects vocabulary is built once in the reference's order, every epoch's shuffle of the pair list is
    drawn on the host (np.random.shuffle, the reference's generator), the steps (train-mode batch-norm network, 1-vs-all
    BCE with label smoothing, Adam over every parameter) run in kp_conve_fit_steps; ExponentialLR is applied between
    epochs exactly as torch chains it."""

    def get_hyperparams_class():
        return BCEOptimizerHyperParams

    def get_kelpie_class():
        return KelpieBCEOptimizer

    @staticmethod
    def er_vocab_tables(rows):
        """extract_er_vocab (:92-96) as arrays: pairs [P, 2] in first-appearance order, CSR of their DISTINCT objects
        (targets[rows, cols] = 1.0 is idempotent, :104)."""
        rows = np.asarray(rows, dtype=np.int64).reshape(-1, 3)
        key = rows[:, 0] * (int(rows[:, 1].max()) + 1 if len(rows) else 1) + rows[:, 1]
        uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
        rank_of_uniq = np.empty(len(uniq), dtype=np.int64)
        by_first = np.argsort(first, kind="stable")
        rank_of_uniq[by_first] = np.arange(len(uniq))
        pid = rank_of_uniq[inv]  # pair id of every row, pairs numbered by first appearance
        pairs = rows[first[by_first]][:, :2]
        po = np.unique(np.stack((pid, rows[:, 2]), 1), axis=0)  # sorted by (pair, object), distinct
        off = np.zeros(len(pairs) + 1, dtype=np.int64)
        np.cumsum(np.bincount(po[:, 0], minlength=len(pairs)), out=off[1:])
        return pairs.astype(np.int32), off, po[:, 1].astype(np.int32)

    def train(self, training_triples, save_path=None, eval_every=-1, valid_triples=None, trial=None, patience=5):
        hp, model = self.hp, self.model
        if isinstance(model, KelpieModel):
            raise Exception("the full-model trainer does not post-train a KelpieModel")
        if not torch.cuda.is_available():
            raise RuntimeError("kelpie_b200 has no CPU training path")
        rows = np.vstack((np.asarray(training_triples), self.dataset.invert_triples(training_triples))).astype(np.int64)
        pairs, pos_off, pos_ids = self.er_vocab_tables(rows)
        n, bs = len(pairs), int(hp["batch_size"])
        model.invalidate_context()  # the scoring context holds copies of the network that is about to change
        model.cuda()
        bn1, bn2, bn3 = model.batch_norm_1, model.batch_norm_2, model.batch_norm_3
        conv, fc = model.convolutional_layer, model.hidden_layer
        held = dict(ent=model.entity_embeddings, rel=model.relation_embeddings, conv_w=conv.weight, conv_b=conv.bias,
                    fc_w=fc.weight, fc_b=fc.bias,
                    bn1_w=bn1.weight, bn1_b=bn1.bias, bn1_mean=bn1.running_mean, bn1_var=bn1.running_var,
                    bn2_w=bn2.weight, bn2_b=bn2.bias, bn2_mean=bn2.running_mean, bn2_var=bn2.running_var,
                    bn3_w=bn3.weight, bn3_b=bn3.bias, bn3_mean=bn3.running_mean, bn3_var=bn3.running_var)
        tensors = {}
        for k, t in held.items():
            t.data = t.data.contiguous().float()
            tensors[k] = t.data
        fit = runtime.ConvEFit(tensors, model.num_filters, model.hidden_layer_size,
                               (model.input_dropout_rate, model.feature_map_dropout_rate, model.hidden_dropout_rate),
                               hp["label_smoothing"], min(bs, n), pairs, pos_off, pos_ids,
                               seed=int(torch.initial_seed()) & 0xFFFFFFFFFFFF)
        off = np.append(np.arange(0, n, bs), n).astype(np.int64)
        order = np.arange(n, dtype=np.int64)
        lr = float(hp["lr"])
        best, bad = None, 0
        self.epoch_losses = []
        try:
            for e in range(1, int(hp["epochs"]) + 1):
                np.random.shuffle(order)  # the same permutation np.random.shuffle applies to the pair list (:114)
                loss = fit.steps(order.astype(np.int32), off, lr, want_loss=self.verbose)
                if loss is not None:
                    self.epoch_losses.append(float(loss.mean()))
                if hp["decay"]:
                    lr = lr * float(hp["decay"])  # ExponentialLR.step(): group["lr"] * gamma (:125-126)
                if valid_triples is not None and eval_every > 0 and e % eval_every == 0:
                    from ..evaluation import Evaluator
                    torch.cuda.synchronize()
                    model.eval()
                    h1 = Evaluator(model).evaluate(valid_triples)["h1"]
                    model.invalidate_context()
                    if trial is not None:
                        trial.report(h1, e)
                        if trial.should_prune():
                            raise RuntimeError("trial pruned")
                    if best is None or h1 > best:
                        best, bad = h1, 0
                    else:
                        bad += 1
                    if bad >= patience:
                        break
            self.launches = fit.launches()
        finally:
            fit.close()
        model.invalidate_context()
        if save_path is not None:
            torch.save(model.state_dict(), save_path)


class KelpieBCEOptimizer(_KelpieOptimizer):
    """bce_optimizer.py:161-208."""
