from .engine import RelevanceEngine
from .post_training_engine import (
    PostTrainingEngine,
    NecessaryPostTrainingEngine,
    SufficientPostTrainingEngine,
)
from .data_poisoning_engine import DPEngine, NecessaryDPEngine, SufficientDPEngine
