"""Mirror of RAGraph_node/ragraph_utils/Propagation.py:5-27."""
import torch

from .. import _lib as L
from ..csr import as_csr


class Propagation:
    @staticmethod
    def aggregate_k_hop_features(adj, x: torch.Tensor, k: int) -> torch.Tensor:
        """k x relu((adj / rowsum(adj)) @ x).  ``adj`` is the dense [n,n] adjacency the reference passes, or a
        CSRGraph / torch sparse tensor.  The row normalisation and ReLU are fused epilogues of the CSR SpMM
        (Propagation.py:15-16,25); k = 0 returns x unchanged."""
        if k <= 0:
            return x
        g = as_csr(adj)
        out = x
        for _ in range(k):
            out = g.spmm(out, L.EPI_ROWNORM | L.EPI_RELU)
        return out
