"""Position-aware structure codes of the node few-shot variant (two-metric retrieval, structure weight 0.001).

Interface parity with RAGraph_node_fewshot/ragraph_utils/PositionAwareEncoder.py:6-48.  SURVEY.md marks the encoder
itself out of the hot-path scope (row 1c: O(n^3) all-pairs shortest paths on query graphs of a few dozen nodes, feeding
a 10-column side input of the similarity kernel), so this is a plain vectorised torch restatement -- no C-ABI kernel --
kept so that ``ToyGraphBase.retrieve(search_keys, search_adj, add_noise)`` of the few-shot variant works with the
reference's three-argument signature.  Differences from the reference are mechanical: the per-(node, anchor) Python
double loop with one ``.item()``-style scalar read per entry becomes one gather + ``torch.where``; the anchors are drawn
with the same CPU ``torch.randint`` call, so a seeded run reproduces the reference's codes bit for bit.
"""
import torch
from torch import Tensor


class PositionAwareEncoder:
    @staticmethod
    def floyd_warshall(adj: Tensor) -> Tensor:
        """All-pairs shortest paths with the adjacency VALUES as edge lengths (0 = no edge), like the reference."""
        n = adj.shape[0]
        dist = adj.clone()
        dist[adj == 0] = float("inf")
        dist.fill_diagonal_(0)
        for k in range(n):
            dist = torch.minimum(dist, dist[:, k].unsqueeze(1) + dist[k, :].unsqueeze(0))
        return dist

    @staticmethod
    def encode_position_aware_code(adj: Tensor, num_anchors: int, dis_q: int = 10) -> Tensor:
        """[n, num_anchors]: 1 / (dist(u, anchor) + 1) where the distance is below dis_q, else 0."""
        if adj.dim() == 3 and adj.shape[0] == 1:
            adj = adj[0]
        distance_matrix = PositionAwareEncoder.floyd_warshall(adj)
        num_nodes = distance_matrix.shape[0]
        anchor_index = torch.randint(low=0, high=num_nodes, size=(int(num_anchors),))       # CPU generator, as the reference
        d = distance_matrix[:, anchor_index.to(distance_matrix.device)]
        return torch.where(d < dis_q, 1 / (d + 1), torch.zeros_like(d))
