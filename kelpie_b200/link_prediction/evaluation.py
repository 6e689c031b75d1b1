"""`Evaluator` (src/link_prediction/evaluation.py:16-92): MRR / H@1 / H@10 / MR from the filtered
ranks of `predict_triples`, which here is one fused-rank launch per direction instead of
batches of 256 with per-row Python filter uploads."""
import numpy as np


class Evaluator:
    def __init__(self, model):
        self.model = model
        self.dataset = model.dataset

    def evaluate(self, triples: np.array, write_output: bool = False, folder: str = "."):
        self.model.eval()
        results = self.model.predict_triples(np.asarray(triples))
        ranks = [r["rank"]["head"] for r in results] + [r["rank"]["tail"] for r in results]
        metrics = {"h1": self.hits_at_k(ranks, 1), "h10": self.hits_at_k(ranks, 10), "mrr": self.mrr(ranks), "mr": self.mr(ranks)}
        if write_output:
            with open(f"{folder}/ranks.csv", "w") as f:
                f.writelines(f"{s};{p};{o};{r['rank']['head']};{r['rank']['tail']}\n"
                             for (s, p, o), r in zip(np.asarray(triples).tolist(), results))
        return metrics

    @staticmethod
    def mrr(values):
        return float(np.mean([1.0 / float(v) for v in values]))

    @staticmethod
    def mr(values):
        return float(np.mean([float(v) for v in values]))

    @staticmethod
    def hits_at_k(values, k: int):
        return float(np.mean([1.0 if v <= k else 0.0 for v in values]))
