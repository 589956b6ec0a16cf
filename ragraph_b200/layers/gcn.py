"""Mirror of RAGraph_node/layers/gcn.py:5-40 (and RAGraph_graph/layers/gcn.py:30-49)."""
import torch
import torch.nn as nn

from .. import _lib as L
from ..csr import as_csr


class GCN(nn.Module):
    def __init__(self, in_ft, out_ft, act=None, bias=True):
        super().__init__()
        self.fc = nn.Linear(in_ft, out_ft, bias=False)
        self.act = nn.PReLU()
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_ft))
        else:
            self.register_parameter('bias', None)
        torch.nn.init.xavier_uniform_(self.fc.weight.data)

    def forward(self, input, sparse=False):
        """PReLU(adj @ (seq @ W^T) + b).  The dense projection stays a library GEMM; the aggregation,
        bias and PReLU are ONE CSR SpMM launch.  ``sparse`` is accepted for signature parity: both the
        dense [n,n] / [1,n,n] adjacency and a torch sparse tensor are converted to CSR."""
        seq, adj = input[0], input[1]
        if seq.dim() == 3 and seq.shape[0] == 1:
            seq = seq[0]
        seq_fts = self.fc(seq)
        epi = L.EPI_PRELU | (L.EPI_BIAS if self.bias is not None else 0)
        # one fused launch under no_grad / detached inputs; with grads enabled the aggregation is the autograd SpMM
        # (dX = A^T dY, same kernel) and bias/PReLU are differentiable torch ops (few-shot decode,
        # RAGraph_node_fewshot/RAGraph.py:69)
        return as_csr(adj).spmm(seq_fts, epi, bias=self.bias, alpha=self.act.weight)
