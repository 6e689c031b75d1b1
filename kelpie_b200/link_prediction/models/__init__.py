from .model import Model, KelpieModel
from .complex import ComplEx, ComplExHyperParams, KelpieComplEx
from .transe import TransE, TransEHyperParams, KelpieTransE
from .conve import ConvE, ConvEHyperParams, KelpieConvE
