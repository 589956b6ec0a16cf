"""Inverse-importance sampling probabilities of the library build, on the CSR kernels (SURVEY.md section 8f rank 2:
"PageRank = SpMV on the same CSR").

Reference: RAGraph_node/ragraph_utils/InverseSampling.py:6-56 (dense [n,n] adjacency, ``torch.mv`` power iteration) and
RAGraph_edge/modules/ragraph_utils/InverseSampling.py:6-69 (torch sparse COO, ``torch.sparse.mm``).  Both compute

    PageRank p  <-  (1 - d)/N + d * (T^T p + dangling/N),   T = row-normalised adjacency, dangling = mass on zero-out-degree rows
    degree centrality = column sums / (N - 1)
    sample_prob ~ 1 / (0.5 * p + 0.5 * centrality + 1e-6), normalised to sum 1

Here the adjacency is converted once to CSR (dense tensor, torch sparse tensor or CSRGraph), the transition matrix is the
same CSR with rescaled values, and every power-iteration step is ONE SpMM launch with F = 1 on its cached transpose; the
dense [n,n] transition matrix of the node variant never exists.  The convergence test keeps the reference's semantics
(L1 change < eps, one host sync per step, the PREVIOUS iterate is returned).  What is done with the probabilities
(``torch.multinomial``, Bernoulli node drop) is RNG-driven build policy and stays with the caller.
"""
from __future__ import annotations

import torch
from torch import Tensor

from .csr import CSRGraph, as_csr


class InverseSampling:
    @staticmethod
    def _values(g: CSRGraph) -> Tensor:
        return g.val if g.val is not None else torch.ones(g.nnz, dtype=torch.float32, device=g.col.device)

    @staticmethod
    def pagerank_algorithm(adj, d: float = 0.85, eps: float = 1e-6, max_iter: int = 100000) -> Tensor:
        g = as_csr(adj)
        n = g.n_rows
        val, rows = InverseSampling._values(g), g.row_ids()
        out_degree = torch.zeros(n, dtype=torch.float32, device=val.device).index_add_(0, rows, val)
        zero_out_degree = out_degree == 0
        out_degree_adj = torch.where(zero_out_degree, torch.ones_like(out_degree), out_degree)
        transition_t = CSRGraph(g.rowptr, g.col, val / out_degree_adj[rows], n, g.n_cols).transpose()
        p = torch.ones(n, dtype=torch.float32, device=val.device) / n
        for _ in range(max_iter):
            dangling_contrib = torch.sum(p[zero_out_degree]) / n
            new_p = (1 - d) / n + d * (transition_t.spmm(p.unsqueeze(1)).squeeze(1) + dangling_contrib)
            if torch.norm(new_p - p, p=1) < eps:
                break
            p = new_p
        return p

    @staticmethod
    def degree_centrality_algorithm(adj) -> Tensor:
        g = as_csr(adj)
        val = InverseSampling._values(g)
        degree = torch.zeros(g.n_cols, dtype=torch.float32, device=val.device).index_add_(0, g.col.to(torch.int64), val)
        return degree / (g.n_rows - 1)

    @staticmethod
    def compute_sample_prob(adj) -> Tensor:
        g = as_csr(adj)
        page_rank = InverseSampling.pagerank_algorithm(g)
        degree_centrality = InverseSampling.degree_centrality_algorithm(g)
        node_importance_alpha, node_importance_eps = 0.5, 1e-6
        node_importance = node_importance_alpha * page_rank + (1 - node_importance_alpha) * degree_centrality
        inverse_node_importance = 1 / (node_importance + node_importance_eps)
        return inverse_node_importance / torch.sum(inverse_node_importance)
