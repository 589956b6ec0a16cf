"""Autograd wiring of the CSR SpMM (training through the propagation kernel).

Where the reference needs gradients through the aggregation:
  * edge `_agg` inside `cal_loss` (RAGraph_edge/modules/RAGraph.py:280-283, 335-350): d loss / d all_emb flows back
    through three `gather * w -> scatter_add_` layers (the adjoint of a scatter is a gather, i.e. A^T);
  * few-shot `decode = GCN2(hidden, adj)` (RAGraph_node_fewshot/RAGraph.py:69, layers/gcn.py:32-40).
Edge weights are never learnable on this path (bi-norm degrees and the time softmax carry no parameters), so only
dX = A^T dY is produced; it is the SAME kernel launched on the transposed CSR, which is built once per graph and
cached on the CSRGraph.  The elementwise epilogues (row normalisation, bias, ReLU/PReLU, blends) are applied by
ordinary differentiable torch ops in training mode; under `torch.no_grad()` callers keep the single fused launch.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import _lib as L
from . import ops
from .csr import CSRGraph


class _SpMMFn(torch.autograd.Function):
    """Y = A @ X with A a CSRGraph (values constant); backward dX = A^T @ dY."""

    @staticmethod
    def forward(ctx, x: Tensor, graph: CSRGraph) -> Tensor:
        ctx.graph = graph
        return ops.csr_spmm(graph.rowptr, graph.col, graph.val, x.contiguous())

    @staticmethod
    def backward(ctx, grad_y: Tensor):
        gt = ctx.graph.transpose()
        return ops.csr_spmm(gt.rowptr, gt.col, gt.val, grad_y.contiguous()), None


def spmm(graph: CSRGraph, x: Tensor) -> Tensor:
    """Differentiable `graph @ x` (x: [n_cols, F] float32 CUDA)."""
    if x.requires_grad and torch.is_grad_enabled():
        return _SpMMFn.apply(x, graph)
    return ops.direct(ops.csr_spmm)(graph.rowptr, graph.col, graph.val, x)


def spmm_epilogue(graph: CSRGraph, x: Tensor, epilogue: int = 0, bias: Optional[Tensor] = None,
                  alpha: Optional[Tensor] = None, blend_in: Optional[Tensor] = None, blend_w: float = 0.0,
                  accum_in: Optional[Tensor] = None) -> Tensor:
    """`rag_csr_spmm_f32` semantics (epilogues in the C-ABI order: ROWNORM, BIAS, RELU, PReLU, BLEND, ACCUM) that is
    differentiable w.r.t. x, bias, alpha, blend_in and accum_in when any of them requires grad; otherwise ONE fused
    launch."""
    needs_grad = torch.is_grad_enabled() and any(
        t is not None and t.requires_grad for t in (x, bias, alpha, blend_in, accum_in))
    if not needs_grad:
        return ops.direct(ops.csr_spmm)(graph.rowptr, graph.col, graph.val, x, epilogue, bias=bias, alpha=alpha,
                                        blend_in=blend_in, blend_w=blend_w, accum_in=accum_in)
    g = graph.row_normalized() if epilogue & L.EPI_ROWNORM else graph
    y = _SpMMFn.apply(x, g) if x.requires_grad else ops.csr_spmm(g.rowptr, g.col, g.val, x)
    if epilogue & L.EPI_BIAS:
        y = y + bias
    if epilogue & L.EPI_RELU:
        y = torch.relu(y)
    if epilogue & L.EPI_PRELU:
        y = torch.where(y >= 0, y, alpha.reshape(-1)[0] * y)
    if epilogue & L.EPI_BLEND:
        y = (1.0 - blend_w) * y + blend_w * blend_in
    if epilogue & L.EPI_ACCUM:
        y = y + accum_in
    return y
