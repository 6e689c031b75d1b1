import torch
from pydantic import BaseModel
from torch.nn import Parameter

from .model import KelpieModel, Model


class ComplExHyperParams(BaseModel):
    dimension: int
    init_scale: float


class ComplEx(Model):
    """complex.py:17-141 -- same constructor, attributes and state-dict keys."""

    def __init__(self, dataset, hp: ComplExHyperParams, init_random=True):
        super().__init__(dataset)
        self.name = "ComplEx"
        self.num_entities = dataset.num_entities
        self.num_relations = 2 * dataset.num_relations
        self.dimension = 2 * hp.dimension
        self.real_dimension = hp.dimension
        self.init_scale = hp.init_scale
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        ent = torch.rand(self.num_entities, self.dimension) if init_random else torch.zeros(self.num_entities, self.dimension)
        rel = torch.rand(self.num_relations, self.dimension) if init_random else torch.zeros(self.num_relations, self.dimension)
        self.entity_embeddings = Parameter(ent.to(dev) * (self.init_scale if init_random else 1.0), requires_grad=True)
        self.relation_embeddings = Parameter(rel.to(dev) * (self.init_scale if init_random else 1.0), requires_grad=True)

    def is_minimizer(self):
        return False

    def kelpie_model_class(self):
        return KelpieComplEx

    def get_hyperparams_class():
        return ComplExHyperParams


class KelpieComplEx(KelpieModel):
    """complex.py:144-160: mimic row = init_tensor * init_scale."""

    def __init__(self, dataset, model: ComplEx, init_tensor, rng_device=None):
        super().__init__(dataset, model, init_tensor.clone() * model.init_scale)
