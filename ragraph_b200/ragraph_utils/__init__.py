"""Host-side mirrors of the reference's ``ragraph_utils`` package for the retrieve -> gather -> propagate path.

Each class keeps the reference's name and call signatures and forwards to the C-ABI kernels:
ToyGraphBase (vector store + retrieve), SimilarityFunctions (materialised cosine, API compatibility), Propagation
(k-hop aggregation as CSR SpMM), TaskDecoder (dense MLP head, torch), PositionAwareEncoder (few-shot structure codes,
plain torch: out of the kernel scope), Augmentation (random build-time augmentation, plain torch, same scope note).  The dataset helpers of the reference package (TU loaders, seeding) are out of scope.
"""
from .Augmentation import Augmentation
from .PositionAwareEncoder import PositionAwareEncoder
from .Propagation import Propagation
from .SimilarityFunctions import SimilarityFunctions
from .TaskDecoder import TaskDecoder
from .ToyGraphBase import ToyGraphBase

__all__ = ["Augmentation", "PositionAwareEncoder", "Propagation", "SimilarityFunctions", "TaskDecoder", "ToyGraphBase"]
