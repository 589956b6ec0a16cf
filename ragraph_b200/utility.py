"""Host-side graph preprocessing on the device, emitting CSR directly.

The reference's ``process_tu_dataset`` (RAGraph_node/ragraph_utils/utility.py:30-72) assembles a DENSE block-diagonal
adjacency with ``np.row_stack`` / ``np.column_stack`` per graph (O(n^2) memory and copies per batch), normalises it
with scipy (``normalize_adj``, utils/process.py:208-215) and ships the dense [n,n] matrix to the GPU, where
Propagation / GCN multiply by it.  Here the same normalised adjacency  D^-1/2 (A + I) D^-1/2  is built as CSR from
the edge lists, on whatever device they live on, in O(E); the dense matrix never exists.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from .csr import CSRGraph


def normalized_adjacency_csr(edge_index: Tensor, num_nodes: int, add_self_loops: bool = True) -> CSRGraph:
    """CSR of normalize_adj(A + I) with A = coo_matrix(ones, (edge_index[0], edge_index[1])) (duplicates add up, as
    scipy's todense does).  normalize_adj computes (A D^-1/2)^T D^-1/2, i.e. entry [i,j] = d_i^-1/2 A[j,i] d_j^-1/2
    with d = row sums of A (+ I); values are formed in float64 and rounded once to float32 like the reference
    (utility.py:66-68)."""
    dev = edge_index.device
    src, dst = edge_index[0].to(torch.int64), edge_index[1].to(torch.int64)      # A[src, dst] += 1
    if add_self_loops:
        loops = torch.arange(num_nodes, device=dev)
        src, dst = torch.cat([src, loops]), torch.cat([dst, loops])
    # coalesce duplicates: A[r, c] = multiplicity
    key = src * num_nodes + dst
    uniq, counts = torch.unique(key, return_counts=True)                           # sorted by (row, col)
    r, c = uniq // num_nodes, uniq % num_nodes
    a = counts.to(torch.float64)
    rowsum = torch.zeros(num_nodes, dtype=torch.float64, device=dev).index_add_(0, r, a)
    d_inv_sqrt = rowsum.pow(-0.5)
    d_inv_sqrt[torch.isinf(d_inv_sqrt)] = 0.0
    val = (d_inv_sqrt[c] * a * d_inv_sqrt[r]).to(torch.float32)                    # out[c, r] = d_c A[r, c] d_r
    # CSRGraph.from_coo groups by edges[:,1] (= output row) with column edges[:,0]
    edges = torch.stack([r, c], dim=1)                                             # out row = c, out col = r
    return CSRGraph.from_coo(edges, val, num_nodes, num_nodes, deterministic=True)


def process_graph_batch(xs: Sequence[Tensor], edge_indices: Sequence[Tensor], num_node_attributes: int
                        ) -> Tuple[Tensor, CSRGraph, Tensor]:
    """process_tu_dataset without the dense adjacency: ``xs[g]`` are the per-graph node matrices [n_g, F] (first
    ``num_node_attributes`` columns = features, rest = node-label one-hots, utility.py:34-41), ``edge_indices[g]`` the
    per-graph [2, E_g] edge lists with LOCAL node ids.  Returns (features [n, num_node_attributes], CSR of the
    block-diagonal normalised adjacency, node_labels [n, F - num_node_attributes])."""
    offs, total = [], 0
    for x in xs:
        offs.append(total)
        total += x.shape[0]
    x_all = torch.cat(list(xs), dim=0)
    ei = torch.cat([e.to(torch.int64) + o for e, o in zip(edge_indices, offs)], dim=1)
    features = x_all[:, :num_node_attributes].float().contiguous()
    node_labels = x_all[:, num_node_attributes:].float().contiguous()
    return features, normalized_adjacency_csr(ei.to(x_all.device), total), node_labels


def process_tu_dataset(data, num_node_attributes: int, device: Optional[torch.device] = None
                       ) -> Tuple[Tensor, CSRGraph, Tensor]:
    """Signature of the reference's ``process_tu_dataset(data, num_node_attributes)`` (RAGraph_node/ragraph_utils/
    utility.py:30-72) over ``process_graph_batch``: ``data`` is a torch_geometric ``Batch`` (or anything exposing
    ``num_graphs`` and ``data[g].x`` / ``data[g].edge_index`` per graph).  Returns (features, adjacency, node_labels) like
    the reference, except that the adjacency is the CSR handle every consumer here accepts in place of the dense [n,n]
    tensor, and nothing goes through numpy / scipy.  ``device`` defaults to the current CUDA device when there is one
    (the reference hard-codes ``.cuda()``), else to where ``data`` lives."""
    xs = [data[g].x for g in range(data.num_graphs)]
    eis = [data[g].edge_index for g in range(data.num_graphs)]
    if device is None and torch.cuda.is_available():
        device = torch.device("cuda", torch.cuda.current_device())
    if device is not None:
        xs, eis = [x.to(device) for x in xs], [e.to(device) for e in eis]
    return process_graph_batch(xs, eis, num_node_attributes)
