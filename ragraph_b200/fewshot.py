"""Few-shot helpers on either side of ``RAGraphFewShot.forward`` (they produce its ``mean_fewshot_logits`` argument and
turn its output logits into labels).

Reference: RAGraph_node_fewshot/ragraph_utils/utility.py:75-162 (same helpers in the graph few-shot variant).  The reference
loops over the unique labels in Python with a boolean mask per label and calls ``F.cosine_similarity`` on broadcast
[n, 1, L] x [1, C, L] tensors; here

  class means of the support logits  -> one SpMM over the 0/1 class-indicator CSR (rows = classes), divided by the counts
  cosine of every logit row against the C class means (+ argmax) -> ``rag_prototype_scores_f32`` (raw mode), no [n, C, L] temporary

Forward only, like the rest of the retrieval side (``fewshot_predict_loss`` is a plain torch MSE on gathered rows).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
from torch import Tensor

from . import _lib as L
from . import ops
from .csr import CSRGraph


def fewshot_mean(fewshot_logits: Tensor, fewshot_labels: Tensor) -> Tuple[Tensor, Tensor]:
    """(mean logits per label [U, L], the sorted unique labels [U]) -- utility.py:75-92."""
    labels = fewshot_labels.reshape(-1)
    unique_labels, inverse = torch.unique(labels, return_inverse=True)           # sorted, like Tensor.unique()
    n, U = labels.numel(), unique_labels.numel()
    edges = torch.stack([torch.arange(n, device=labels.device), inverse], dim=1)  # row = class slot, col = support sample
    sums = CSRGraph.from_coo(edges, None, U, n, deterministic=True).spmm(fewshot_logits.float().contiguous())
    counts = torch.bincount(inverse, minlength=U).to(sums.dtype).unsqueeze(1)
    return sums / counts, unique_labels


def fewshot_logits_map(fewshot_logits: Tensor, fewshot_labels: Tensor) -> Dict[int, Tensor]:
    """{label: mean logits} -- utility.py:95-100."""
    mean_fewshot_logits, unique_labels = fewshot_mean(fewshot_logits, fewshot_labels)
    return {int(label): logit for label, logit in zip(unique_labels.tolist(), mean_fewshot_logits)}


def fewshot_mean_logits(fewshot_logits: Tensor, fewshot_labels: Tensor) -> Tensor:
    """[C, L] mean logits ordered by label id 0..C-1 (labels must be exactly 0..C-1, as the reference's dict lookup
    ``label_to_logit[label] for label in range(len(...))`` requires) -- utility.py:115-127."""
    mean_fewshot_logits, unique_labels = fewshot_mean(fewshot_logits, fewshot_labels)
    expect = torch.arange(unique_labels.numel(), device=unique_labels.device, dtype=unique_labels.dtype)
    if not torch.equal(unique_labels, expect):
        raise KeyError(f"fewshot_mean_logits: support labels {unique_labels.tolist()} are not 0..{unique_labels.numel() - 1}")
    return mean_fewshot_logits


def fewshot_predict_logits(mean_fewshot_logits: Tensor, logits: Tensor) -> Tensor:
    """[n, C] cosine similarity of every logit row with every class mean -- utility.py:129-134."""
    return ops.prototype_scores(logits.float(), mean_fewshot_logits.float(), L.SCORES_RAW)


def fewshot_predict_labels_by_mean(mean_fewshot_logits: Tensor, logits: Tensor) -> Tensor:
    """index of the most similar class mean per row -- utility.py:152-162."""
    return torch.argmax(fewshot_predict_logits(mean_fewshot_logits, logits), dim=1)


def fewshot_predict_labels(fewshot_logits: Tensor, fewshot_labels: Tensor, logits: Tensor) -> Tensor:
    """label (not index) of the most similar class mean per row -- utility.py:136-150."""
    mean_fewshot_logits, unique_labels = fewshot_mean(fewshot_logits, fewshot_labels)
    return unique_labels[fewshot_predict_labels_by_mean(mean_fewshot_logits, logits)]


def fewshot_predict_loss(fewshot_logits: Tensor, fewshot_labels: Tensor, logits: Tensor, labels: Tensor) -> Tensor:
    """MSE between ``logits`` and the mean support logits of each row's gold label -- utility.py:102-113."""
    mean_fewshot_logits, unique_labels = fewshot_mean(fewshot_logits, fewshot_labels)
    slot = torch.searchsorted(unique_labels, labels.reshape(-1).to(unique_labels.dtype))
    if bool((slot >= unique_labels.numel()).any()) or not torch.equal(unique_labels[slot], labels.reshape(-1).to(unique_labels.dtype)):
        raise KeyError("fewshot_predict_loss: a gold label does not occur among the support labels")
    gold_logits = ops.direct(ops.gather_rows)(mean_fewshot_logits.contiguous(), slot)
    return torch.nn.functional.mse_loss(logits, gold_logits.to(logits.device))
