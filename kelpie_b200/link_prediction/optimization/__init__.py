from .optimizers import (
    Optimizer,
    PairwiseRankingOptimizer, KelpiePairwiseRankingOptimizer, PairwiseRankingOptimizerHyperParams,
    MultiClassNLLOptimizer, KelpieMultiClassNLLOptimizer, MultiClassNLLOptimizerHyperParams,
    BCEOptimizer, KelpieBCEOptimizer, BCEOptimizerHyperParams,
)
