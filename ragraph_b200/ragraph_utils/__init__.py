"""Host-side mirrors of the reference's ``ragraph_utils`` package for the retrieve -> gather -> propagate path.

Each class keeps the reference's name and call signatures and forwards to the C-ABI kernels:
ToyGraphBase (vector store + retrieve), SimilarityFunctions (materialised cosine, API compatibility), Propagation
(k-hop aggregation as CSR SpMM), TaskDecoder (dense MLP head, torch), PositionAwareEncoder (few-shot structure codes,
plain torch: out of the kernel scope).  Library-construction policy (augmentation,
inverse sampling draws) and the dataset helpers of the reference package are out of scope.
"""
from .PositionAwareEncoder import PositionAwareEncoder
from .Propagation import Propagation
from .SimilarityFunctions import SimilarityFunctions
from .TaskDecoder import TaskDecoder
from .ToyGraphBase import ToyGraphBase

__all__ = ["PositionAwareEncoder", "Propagation", "SimilarityFunctions", "TaskDecoder", "ToyGraphBase"]
