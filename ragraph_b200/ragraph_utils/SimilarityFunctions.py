"""Mirror of RAGraph_node/ragraph_utils/SimilarityFunctions.py:5-16 (identical in all five variants)."""
import torch

from .. import ops


class SimilarityFunctions:
    @staticmethod
    def calculate_cosine_similarity(search_keys: torch.Tensor, resource_keys: torch.Tensor) -> torch.Tensor:
        """[Q,d] x [N,d] -> [Q,N] cosine similarities (both sides L2-normalised with the 1e-12 clamp).
        A 1-D query [d] (graph variant, RAGraph_graph/RAGraph.py:50) returns [N].  Materialises the matrix,
        as the reference API demands; retrieval itself uses the fused ``ops.cosine_topk``."""
        one_d = search_keys.dim() == 1
        q = search_keys.unsqueeze(0) if one_d else search_keys
        out = ops.cosine_similarity(q, resource_keys)
        return out[0] if one_d else out
