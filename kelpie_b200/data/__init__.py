from .dataset import Dataset
from .kelpie_dataset import KelpieDataset
from .names import *
