from .engine import RelevanceEngine
from .post_training_engine import (
    PostTrainingEngine,
    NecessaryPostTrainingEngine,
    SufficientPostTrainingEngine,
)
