"""Mirror of the reference's downstream prompt.

node : RAGraph_node/downprompt.py:6-48 (downprompt), :59-79 (averageemb), :118-130 (downstreamprompt)
graph: RAGraph_graph/downprompt.py:6-31 (downprompt), :41-56 (predict), :59-94 (averageemb),
       :98-112 (split_and_batchify_graph_feats), :197-209 (downstreamprompt)

What the reference does in Python loops with one ``.item()`` / scalar ``cosine_similarity`` per (row, class) runs
here as single launches of the C-ABI kernels:

  ELU(weight * emb) / weight * emb                -> rag_prompt_act_f32
  cosine vs the C class prototypes + (log_)softmax -> rag_prototype_scores_f32 (prompt applied on the fly, the
                                                     [n,d] prompted copy is not written when only scores are needed)
  class-prototype means, per-graph readout sums   -> rag_csr_spmm_f32 over a 0/1 indicator CSR (rows = classes or
                                                     graphs), no Python loop over nodes

Semantics kept, quirks included: the node variant averages class sums over ``n // 2`` slots and the graph variant
over ``n`` slots (the reference takes ``torch.mean`` over a scratch tensor of that many rows per class and relies on
the unwritten rows being zero -- it allocates them uninitialised; zero is the only value for which its result is
defined, and is what we implement).  Forward only: like every other retrieval-side op the prompt carries no autograd
formula here (the reference never imports downprompt from its RAG scripts; it is a GraphPrompt baseline head).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib as L
from . import ops
from .csr import CSRGraph


def _indicator_csr_by_label(labels: Tensor, nb_class: int, n: int) -> CSRGraph:
    """CSR [nb_class, n] with a 1 at (labels[i], i); labels outside [0, nb_class) own no row (the reference's
    chain of ``if labels[x].item() == c`` ignores them too).  Rows keep node order, so sums run in node order."""
    lab = labels.reshape(-1).to(torch.int64)
    node = torch.arange(n, device=lab.device)
    keep = (lab >= 0) & (lab < nb_class)
    if not bool(keep.all()):
        node, lab = node[keep], lab[keep]
    edges = torch.stack([node, lab], dim=1)
    return CSRGraph.from_coo(edges, None, nb_class, n, deterministic=True)


def averageemb(labels: Tensor, rawret: Tensor, nb_class: int = 3, slots: Optional[int] = None) -> Tensor:
    """Class prototypes [nb_class, d] = (sum of the rows of each class) / slots.
    node (RAGraph_node/downprompt.py:59-79): nb_class = 3, slots = n // 2 (the default here);
    graph (RAGraph_graph/downprompt.py:59-94): slots = n (pass ``slots=rawret.shape[0]``)."""
    n = rawret.shape[0]
    slots = int(n / 2) if slots is None else int(slots)
    sums = _indicator_csr_by_label(labels, nb_class, n).spmm(rawret.contiguous())
    return sums / slots


def split_and_batchify_graph_feats(batched_graph_feats: Tensor, graph_sizes: Tensor) -> Tensor:
    """Per-graph sum readout (RAGraph_graph/downprompt.py:98-112): graph i owns the next graph_sizes[i] rows."""
    sizes = graph_sizes.reshape(-1).to(torch.int64)
    G, n = sizes.numel(), batched_graph_feats.shape[0]
    rowptr = torch.zeros(G + 1, dtype=torch.int64, device=batched_graph_feats.device)
    torch.cumsum(sizes.to(rowptr.device), 0, out=rowptr[1:])
    col = torch.arange(n, dtype=torch.int32, device=batched_graph_feats.device)
    nnz = int(rowptr[-1].item())
    if nnz > n:
        raise RuntimeError(f"split_and_batchify_graph_feats: graph sizes sum to {nnz} > {n} rows")
    return CSRGraph(rowptr, col[:nnz], None, G, n).spmm(batched_graph_feats.contiguous())


def predict(graphnum: int, nb_classes: int, rawret: Tensor, ave: Tensor) -> Tensor:
    """log_softmax over the cosine of each graph embedding against the class prototypes
    (RAGraph_graph/downprompt.py:41-56; the reference fills columns for nb_classes 2 or 6 only, any C <= 32 here)."""
    return ops.prototype_scores(rawret[:graphnum], ave[:nb_classes], L.SCORES_LOG_SOFTMAX)


class downstreamprompt(nn.Module):
    """weight [1, hid] (xavier) times the embedding; ELU in the node variant, identity in the graph variant."""

    def __init__(self, hid_units: int, variant: str = "node"):
        super().__init__()
        assert variant in ("node", "graph")
        self.variant = variant
        self.weight = nn.Parameter(torch.empty(1, hid_units), requires_grad=False)
        self.reset_parameters()

    @property
    def act(self) -> int:
        return L.ACT_ELU if self.variant == "node" else L.ACT_NONE

    def reset_parameters(self):
        torch.nn.init.xavier_uniform_(self.weight)

    def forward(self, graph_embedding: Tensor) -> Tensor:
        return ops.prompt_act(graph_embedding, self.weight, self.act)


class downprompt(nn.Module):
    """node : downprompt(prompt1, prompt2, prompt3, ft_in, nb_classes, feature, labels); forward(seq, train=0)
              -> class probabilities [n, C] (softmax of the cosine against the class prototypes)
       graph: downprompt(prompt1, prompt2, prompt3, ft_in, nb_classes); forward(seq, graph_len)
              -> prompted per-graph sum readout [G, ft_in]"""

    def __init__(self, prompt1, prompt2, prompt3, ft_in, nb_classes, feature: Optional[Tensor] = None,
                 labels: Optional[Tensor] = None):
        super().__init__()
        self.variant = "node" if feature is not None else "graph"
        self.nb_classes = nb_classes
        self.labels = labels
        self.downprompt = downstreamprompt(ft_in, self.variant)
        self.prompt = torch.cat((prompt1, prompt2, prompt3), 0)
        if self.variant == "node":
            feature = feature.squeeze()
            self.ave = averageemb(labels=self.labels, rawret=feature, nb_class=nb_classes)

    def forward(self, seq: Tensor, arg=0) -> Tensor:
        if self.variant == "graph":
            return split_and_batchify_graph_feats(self.downprompt(seq), arg)
        if arg == 1:                                  # train: prototypes follow the prompted embeddings (:32-33)
            rawret = self.downprompt(seq)
            self.ave = averageemb(labels=self.labels, rawret=rawret, nb_class=self.nb_classes)
            return ops.prototype_scores(rawret, self.ave, L.SCORES_SOFTMAX)
        return ops.prototype_scores(seq, self.ave, L.SCORES_SOFTMAX, self.downprompt.weight, self.downprompt.act)
