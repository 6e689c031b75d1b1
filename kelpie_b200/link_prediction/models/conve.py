import torch
from pydantic import BaseModel
from torch import nn
from torch.nn import Parameter
from torch.nn.init import xavier_normal_

from ... import plans
from .model import KelpieModel, Model


class ConvEHyperParams(BaseModel):
    dimension: int
    input_dropout_rate: float
    feature_map_dropout_rate: float
    hidden_dropout_rate: float
    hidden_layer_size: int


class ConvE(Model):
    """conve.py:23-190 -- same constructor, sub-module names and state-dict keys."""

    def __init__(self, dataset, hp: ConvEHyperParams, init_random=True):
        super().__init__(dataset)
        self.name = "ConvE"
        self.num_entities = dataset.num_entities
        self.num_relations = 2 * dataset.num_relations
        self.dimension = hp.dimension
        self.input_dropout_rate = hp.input_dropout_rate
        self.feature_map_dropout_rate = hp.feature_map_dropout_rate
        self.hidden_dropout_rate = hp.hidden_dropout_rate
        self.hidden_layer_size = hp.hidden_layer_size
        self.embedding_width = 20
        self.embedding_height = self.dimension // self.embedding_width
        self.kernel_shape = (3, 3)
        self.num_filters = 32
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        self.input_dropout = nn.Dropout(self.input_dropout_rate)
        self.feature_map_dropout = nn.Dropout2d(self.feature_map_dropout_rate)
        self.hidden_dropout = nn.Dropout(self.hidden_dropout_rate)
        self.batch_norm_1 = nn.BatchNorm2d(1).to(dev)
        self.batch_norm_2 = nn.BatchNorm2d(self.num_filters).to(dev)
        self.batch_norm_3 = nn.BatchNorm1d(self.dimension).to(dev)
        self.convolutional_layer = nn.Conv2d(1, self.num_filters, self.kernel_shape, 1, 0, bias=True).to(dev)
        self.hidden_layer = nn.Linear(self.hidden_layer_size, self.dimension).to(dev)
        ent = torch.rand(self.num_entities, self.dimension) if init_random else torch.zeros(self.num_entities, self.dimension)
        rel = torch.rand(self.num_relations, self.dimension) if init_random else torch.zeros(self.num_relations, self.dimension)
        self.entity_embeddings = Parameter(ent.to(dev), requires_grad=True)
        self.relation_embeddings = Parameter(rel.to(dev), requires_grad=True)
        if init_random:
            xavier_normal_(self.entity_embeddings)
            xavier_normal_(self.relation_embeddings)

    def is_minimizer(self):
        return False

    def kelpie_model_class(self):
        return KelpieConvE

    def get_hyperparams_class():
        return ConvEHyperParams


_BURN_CHECKED = {}


def burn_conve_constructor_rng(model):
    """KelpieConvE builds a throw-away ConvE whose nn.Conv2d / nn.Linear initialisers draw from
    the CPU generator (conve.py:202 -> :50-52) before being replaced by deep copies; replaying
    those draws keeps every later random number aligned with the reference."""
    F, h, d = int(model.num_filters), int(model.hidden_layer_size), int(model.dimension)
    words = 9 * F + F + h * d + d  # one generator word per fp32 element of the two weights and two biases
    ok = _BURN_CHECKED.get((F, h, d))
    if ok is None:  # once per shape: the native skip must leave the generator where the constructors leave it
        ok = False
        if plans.HostReplay.available():
            start = torch.get_rng_state()
            plans.HostReplay.torch_skip(words)
            mine = torch.rand(4)
            torch.set_rng_state(start)
            nn.Conv2d(1, F, (3, 3), 1, 0, bias=True)
            nn.Linear(h, d)
            ok = bool(torch.equal(mine, torch.rand(4)))
            torch.set_rng_state(start)
        _BURN_CHECKED[(F, h, d)] = ok
    if ok:
        plans.HostReplay.torch_skip(words)  # 0.9 ms instead of 13 ms of discarded kaiming_uniform_ draws
        return
    nn.Conv2d(1, F, (3, 3), 1, 0, bias=True)
    nn.Linear(h, d)


class KelpieConvE(KelpieModel):
    """conve.py:193-237: mimic row = init_tensor as is; the frozen network is shared."""

    def __init__(self, dataset, model: ConvE, init_tensor, rng_device=None, replay_rng=True):
        if replay_rng:
            burn_conve_constructor_rng(model)
        super().__init__(dataset, model, init_tensor.clone())
