"""Random augmentation of resource graphs at library-build time (policy, not a kernel).

Interface parity with RAGraph_node/ragraph_utils/Augmentation.py:5-64 (identical in every variant).  SURVEY.md marks it
out of the hot-path scope (row 1a: RNG-driven elementwise work on graphs of a few dozen nodes, executed once per library
build); it is shipped as plain torch so that ``ToyGraphBase._build_toy_graph_base`` can be run end to end.  The random
draws are issued in the reference's order with the reference's shapes -- Gaussian feature noise, then the Bernoulli node
mask, then one uniform matrix for the edge rewrite -- so a seeded build reproduces the reference's library.  The sampling
probabilities come from ``ragraph_b200.sampling.InverseSampling`` (PageRank as SpMV steps on the CSR kernel).

Kept as in the reference, on purpose: the node mask is drawn with probability ``sample_prob * 0.01`` (about 1/(100 n) per
node), so augmented feature matrices are almost entirely zero, and the rewritten adjacency is a 0/1 matrix with edge
probability (p_i + p_j) / 2 that is NOT re-normalised.
"""
import torch
from torch import Tensor

from ..sampling import InverseSampling

FEATURE_NOISE_STD = 0.1
NODE_KEEP_RATE = 0.01


class Augmentation:
    @staticmethod
    def augment_features(features: Tensor, sample_prob: Tensor) -> Tensor:
        noisy = features + FEATURE_NOISE_STD * torch.randn_like(features)
        keep = torch.bernoulli(NODE_KEEP_RATE * sample_prob)
        return noisy * keep.unsqueeze(-1)

    @staticmethod
    def augment_adj(adj: Tensor, sample_prob: Tensor) -> Tensor:
        edge_prob = 0.5 * (sample_prob.unsqueeze(1) + sample_prob.unsqueeze(0))
        draws = torch.rand(adj.shape, device=adj.device)
        return (draws < edge_prob).to(adj.dtype)

    @staticmethod
    def augment_graph(num_augment_scale: int, features: Tensor, adj: Tensor):
        """Yields the original (features, adj) first, then ``num_augment_scale`` augmented copies."""
        sample_prob = InverseSampling.compute_sample_prob(adj)
        yield features, adj
        for _ in range(int(num_augment_scale)):
            yield Augmentation.augment_features(features, sample_prob), Augmentation.augment_adj(adj, sample_prob)
