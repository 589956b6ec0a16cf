from .TaskDecoder import TaskDecoder
from .ToyGraphBase import ToyGraphBase
from .Propagation import Propagation
from .SimilarityFunctions import SimilarityFunctions

__all__ = ["TaskDecoder", "ToyGraphBase", "Propagation", "SimilarityFunctions"]
