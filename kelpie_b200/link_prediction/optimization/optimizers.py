"""`Optimizer(model, hp, verbose).train(training_triples)` interface of the reference
(src/link_prediction/optimization/*.py).  Only the Kelpie* subclasses -- the mimic
post-training -- are on the hot path; they run as ONE job through the batched CUDA kernels
(the engines batch many).  The full-model trainers stay with the reference.
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
        rows = ctx.post_train(runtime.make_hp(base.name, self.hp), **batch.arrays())
        with torch.no_grad():
            self.model.kelpie_entity_emb = rows[:1].clone()
        self.model.update_embeddings()


class PairwiseRankingOptimizer(Optimizer):
    def get_hyperparams_class():
        return PairwiseRankingOptimizerHyperParams

    def get_kelpie_class():
        return KelpiePairwiseRankingOptimizer


class KelpiePairwiseRankingOptimizer(_KelpieOptimizer):
    """pairwise_ranking_optimizer.py:160-203."""


class MultiClassNLLOptimizer(Optimizer):
    def get_hyperparams_class():
        return MultiClassNLLOptimizerHyperParams

    def get_kelpie_class():
        return KelpieMultiClassNLLOptimizer


class KelpieMultiClassNLLOptimizer(_KelpieOptimizer):
    """multiclass_nll_optimizer.py:138-164."""


class BCEOptimizer(Optimizer):
    def get_hyperparams_class():
        return BCEOptimizerHyperParams

    def get_kelpie_class():
        return KelpieBCEOptimizer


class KelpieBCEOptimizer(_KelpieOptimizer):
    """bce_optimizer.py:161-208."""
