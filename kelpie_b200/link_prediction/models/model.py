"""`Model` / `KelpieModel` interfaces of the reference (src/link_prediction/models/model.py:8-128)
with score / predict / rank running in the CUDA library instead of torch ops."""
import numpy as np
import torch
from torch import nn

from ... import runtime


def model_weights(model):
    """Frozen weights of a reference-style model object (ours or the reference's own class)."""
    kind = model.name
    w = dict(kind=kind, ent=model.entity_embeddings.detach(), rel=model.relation_embeddings.detach(), norm=2, conve=None)
    if kind == "TransE":
        w["norm"] = int(model.norm)
    if kind == "ConvE":
        bn = lambda m, i: {f"bn{i}_w": m.weight.detach(), f"bn{i}_b": m.bias.detach(),
                           f"bn{i}_mean": m.running_mean.detach(), f"bn{i}_var": m.running_var.detach()}
        c = dict(conv_w=model.convolutional_layer.weight.detach(), conv_b=model.convolutional_layer.bias.detach(),
                 fc_w=model.hidden_layer.weight.detach(), fc_b=model.hidden_layer.bias.detach())
        c.update(bn(model.batch_norm_1, 1))
        c.update(bn(model.batch_norm_2, 2))
        c.update(bn(model.batch_norm_3, 3))
        c["dropout"] = (float(model.input_dropout_rate), float(model.feature_map_dropout_rate), float(model.hidden_dropout_rate))
        w["conve"] = c
    return w


def context_for(model):
    """The (cached) device context of a model: tables + resident filter CSR of its dataset."""
    ctx = getattr(model, "_kp_ctx", None)
    if ctx is None:
        w = model_weights(model)
        ctx = runtime.Context(w["kind"], w["ent"], w["rel"], norm=w["norm"], conve=w["conve"])
        ds = model.dataset
        if hasattr(ds, "filter_facts") and ctx.N * ctx.R2 < 2 ** 31:
            ctx.build_filter(ds.filter_facts())  # sort / unique / segment on the device (kp_filter_build)
        else:
            ctx.upload_filter(ds.to_filter)
        object.__setattr__(model, "_kp_ctx", ctx)
    return ctx


class Model(nn.Module):
    def __init__(self, dataset):
        super().__init__()
        self.dataset = dataset

    def is_minimizer(self):
        pass

    def context(self):
        return context_for(self)

    def invalidate_context(self):
        """Call after changing the embeddings in place (the device context borrows them)."""
        ctx = getattr(self, "_kp_ctx", None)
        if ctx is not None:
            ctx.close()
            object.__setattr__(self, "_kp_ctx", None)

    def all_scores(self, triples):
        """[Q,3] -> [Q, N] fp32 cuda tensor (transe.py:48-65, complex.py:88-113, conve.py:133-158)."""
        return self.context().all_scores(np.asarray(triples))

    def score(self, triples):
        """transe.py:38-46, complex.py:41-56, conve.py:68-75: the score of each triple itself -- one row of work per
        triple (kp_score_triples), not a pass over the entity table."""
        sc = self.context().score_triples(np.asarray(triples)).cpu().numpy()
        return sc.reshape(-1, 1) if self.name == "ComplEx" else sc  # complex.py:56 keeps the summed dimension

    def forward(self, triples):
        """transe.py:67-75 / complex.py:59-86 / conve.py:65-66: (scores, regulariser factors) as the reference's
        optimizers consume them.  Scores come from the device kernels; the factors are row gathers of the tables."""
        triples = np.asarray(triples)
        if self.name == "ConvE":
            return self.all_scores(triples)
        dev = self.entity_embeddings.device
        t = torch.as_tensor(triples, device=dev, dtype=torch.long)
        lhs, rel, rhs = self.entity_embeddings[t[:, 0]], self.relation_embeddings[t[:, 1]], self.entity_embeddings[t[:, 2]]
        if self.name == "TransE":
            return self.context().score_triples(triples), (lhs, rel, rhs)
        d = self.real_dimension
        factors = tuple(torch.sqrt(x[:, :d] ** 2 + x[:, d:] ** 2) for x in (lhs, rel, rhs))
        return self.all_scores(triples), factors

    def predict_tails(self, triples):
        """model.py:42-68 / conve.py:160-184: (target scores, filtered tail ranks)."""
        triples = np.asarray(triples)
        mode = runtime.RANK_CONVE_SORT if self.name == "ConvE" else runtime.RANK_MODEL
        ts, _, rk = self.context().filtered_rank(triples, mode)
        scores = [x for x in ts.cpu().numpy()]
        ranks = rk.cpu().numpy()
        ranks = [int(r) for r in ranks] if self.name == "ConvE" else [float(r) for r in ranks]
        return scores, ranks

    def predict_triples(self, triples):
        """model.py:25-40."""
        direct = np.asarray(triples)
        assert np.all(direct[:, 1] < self.dataset.num_relations)
        ds_, tr = self.predict_tails(direct)
        hs_, hr = self.predict_tails(self.dataset.invert_triples(direct))
        return [{"score": {"tail": ds_[i], "head": hs_[i]}, "rank": {"tail": int(tr[i]), "head": int(hr[i])}}
                for i in range(direct.shape[0])]

    def kelpie_model_class(self):
        pass


class KelpieModel:
    """Mimic wrapper (model.py:80-128): the frozen tables stay in the base model's device
    context (never cloned); the mimic row is the only state."""

    def __init__(self, dataset, model, init_row):
        self.model = model
        self._dataset = dataset
        self.original_entity = dataset.original_entity
        self.kelpie_entity = dataset.kelpie_entity
        self.kelpie_entity_emb = init_row
        self.training = False

    @property
    def name(self):
        return self.model.name

    @property
    def dimension(self):
        return self.model.dimension

    @property
    def entity_embeddings(self):
        return torch.cat([self.model.entity_embeddings.detach(), self.kelpie_entity_emb.to(self.model.entity_embeddings.device)], 0)

    @property
    def dataset(self):
        return self._dataset

    def to(self, device):
        return self

    def cuda(self):
        return self

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        self.training = mode
        return self

    def is_minimizer(self):
        return self.model.is_minimizer()

    def parameters(self):
        return [self.kelpie_entity_emb]

    def update_embeddings(self):
        pass  # the mimic row is passed to the kernels directly; no table row to refresh

    def all_scores(self, triples):
        """[Q,3] (entity id N = the mimic) -> [Q, N+1] fp32 cuda tensor."""
        triples = np.asarray(triples)
        ctx = context_for(self.model)
        rows = self.kelpie_entity_emb.detach().to(ctx.device).view(1, -1).expand(len(triples), -1).contiguous()
        return ctx.all_scores(triples, mimic_rows=rows)

    def kelpie_model_class(self):
        raise Exception(self.__class__.__name__ + " is a KelpieModel.")
