"""kelpie_b200 -- B200-native relevance-engine hot path of Kelpie++ (rbarile17/kelpie).

Mimic post-training and the score-all-entities filtered rank for TransE / ComplEx / ConvE as
hand-written CUDA for sm_100a behind the reference's Python interfaces
(`relevance_engines`, `link_prediction`, `data`).  See DESIGN.md / INTEGRATION.md.
"""
MODELS_PATH = "models"
RESULTS_PATH = "results"
key = lambda x: x[1]  # src/__init__.py:38

__version__ = "0.1.0"
