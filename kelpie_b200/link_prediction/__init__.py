from .optimization import BCEOptimizer, MultiClassNLLOptimizer, PairwiseRankingOptimizer
from .models import ConvE, ComplEx, TransE

# src/link_prediction/__init__.py:5-9
MODEL_REGISTRY = {
    "ComplEx": {"class": ComplEx, "optimizer": MultiClassNLLOptimizer},
    "TransE": {"class": TransE, "optimizer": PairwiseRankingOptimizer},
    "ConvE": {"class": ConvE, "optimizer": BCEOptimizer},
}
