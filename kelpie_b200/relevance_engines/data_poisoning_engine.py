"""`DPEngine` / `NecessaryDPEngine` / `SufficientDPEngine` (src/relevance_engines/data_poisoning_engine.py:9-141, SURVEY
8f-4): the data-poisoning baseline.  Same constructor and `compute_relevance(pred, perspective, triple)`; the score /
gradient / perturbed-score arithmetic runs in kp_dp_relevance, and `compute_relevances` evaluates many facts of one
prediction in one launch.  Like the reference it needs `Model.score_embeddings`, i.e. ComplEx among the in-scope models
(TransE and ConvE define `score_embs` only, so the reference raises AttributeError for them; here: RuntimeError)."""
from ..data import Dataset
from ..link_prediction.models.model import context_for
from .engine import RelevanceEngine


class DPEngine(RelevanceEngine):
    def __init__(self, model, dataset, epsilon: float):
        RelevanceEngine.__init__(self, model=model, dataset=dataset)
        self.epsilon = epsilon
        self.lambd = 1

    sufficient = False

    def _launch(self, jobs):
        """jobs: [(pred, triple, entity)] -> list of python floats (float32 values, as scores[0] - scores[1] gives)."""
        if not jobs:
            return []
        out = context_for(self.model).dp_relevance([j[0] for j in jobs], [j[1] for j in jobs], [j[2] for j in jobs],
                                                   self.epsilon, self.lambd, self.sufficient)
        return [x for x in out.cpu().numpy()]

    def compute_relevance(self, pred, perspective: str, triple):
        raise NotImplementedError


class NecessaryDPEngine(DPEngine):
    def compute_relevances(self, pred, perspective: str, triples):
        pred_s, _, pred_o = pred
        entity = pred_s if perspective == "head" else pred_o
        return self._launch([(tuple(pred), tuple(t), entity) for t in triples])

    def compute_relevance(self, pred, perspective: str, triple):
        return self.compute_relevances(pred, perspective, [triple])[0]


class SufficientDPEngine(DPEngine):
    sufficient = True

    def _jobs(self, pred, perspective, triple):
        """data_poisoning_engine.py:133-141, including its quirk: `triple` and `pred` are REASSIGNED inside the loop while
        pred_s keeps the original head, so from the second conversion entity on nothing is left to replace."""
        pred_s = pred[0]
        jobs = []
        for entity in self.entities_to_convert:
            triple = Dataset.replace_entity_in_triple(triple, pred_s, entity)
            pred = Dataset.replace_entity_in_triple(pred, pred_s, entity)
            jobs.append((tuple(pred), tuple(triple), pred[0] if perspective == "head" else pred[2]))
        return jobs

    def compute_individual_relevance(self, pred, perspective: str, triple):
        entity = pred[0] if perspective == "head" else pred[2]
        return self._launch([(tuple(pred), tuple(triple), entity)])[0]

    def compute_relevances(self, pred, perspective: str, triples):
        per = [self._jobs(pred, perspective, t) for t in triples]
        flat = self._launch([j for jobs in per for j in jobs])
        out, k = [], 0
        for jobs in per:
            vals = flat[k:k + len(jobs)]
            k += len(jobs)
            out.append(sum(vals) / len(vals))
        return out

    def compute_relevance(self, pred, perspective: str, triple):
        return self.compute_relevances(pred, perspective, [triple])[0]
