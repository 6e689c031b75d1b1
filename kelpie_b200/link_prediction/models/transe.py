import torch
from pydantic import BaseModel
from torch.nn import Parameter
from torch.nn.init import xavier_normal_

from .model import KelpieModel, Model


class TransEHyperParams(BaseModel):
    dimension: int
    norm: int


class TransE(Model):
    """transe.py:17-81 -- same constructor, attributes and state-dict keys."""

    def __init__(self, dataset, hp: TransEHyperParams, init_random=True):
        super().__init__(dataset)
        self.name = "TransE"
        self.num_entities = dataset.num_entities
        self.num_relations = 2 * dataset.num_relations
        self.dimension = hp.dimension
        self.norm = hp.norm
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        ent = torch.rand(self.num_entities, self.dimension) if init_random else torch.zeros(self.num_entities, self.dimension)
        rel = torch.rand(self.num_relations, self.dimension) if init_random else torch.zeros(self.num_relations, self.dimension)
        self.entity_embeddings = Parameter(ent.to(dev), requires_grad=True)
        self.relation_embeddings = Parameter(rel.to(dev), requires_grad=True)
        if init_random:
            xavier_normal_(self.entity_embeddings)
            xavier_normal_(self.relation_embeddings)

    def is_minimizer(self):
        return True

    def kelpie_model_class(self):
        return KelpieTransE

    def get_hyperparams_class():
        return TransEHyperParams


class KelpieTransE(KelpieModel):
    """transe.py:84-99: the mimic row is re-drawn with xavier_normal_ on its own device."""

    def __init__(self, dataset, model: TransE, init_tensor, rng_device=None):
        dev = rng_device or model.entity_embeddings.device
        row = init_tensor.clone().to(dev)
        xavier_normal_(row)
        super().__init__(dataset, model, row)
