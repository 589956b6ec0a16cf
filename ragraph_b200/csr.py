"""CSR handle used by the propagation ops (what the dense block-diagonal adjacency of the
node/graph variants and the COO edge list of the edge variant are converted to, once)."""
from __future__ import annotations

import weakref
from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor

from . import ops


@dataclass
class CSRGraph:
    rowptr: Tensor            # int64 [n_rows+1]
    col: Tensor               # int32 [nnz]
    val: Optional[Tensor]     # float32 [nnz] or None (= ones)
    n_rows: int
    n_cols: int

    @property
    def nnz(self) -> int:
        return self.col.numel()

    @staticmethod
    def from_dense(adj: Tensor) -> "CSRGraph":
        """Dense adjacency ([n,n] or [1,n,n] as layers/gcn.py:36 squeezes it) -> CSR.  Cached per tensor
        object + version so RAGraph.forward, which hands the same adj to GCN and Propagation, converts once."""
        if adj.dim() == 3 and adj.shape[0] == 1:
            adj = adj[0]
        key = id(adj)
        hit = _dense_cache.get(key)
        if hit is not None and hit[0]() is adj and hit[1] == adj._version:
            return hit[2]
        rowptr, col, val = ops.csr_from_dense(adj)
        g = CSRGraph(rowptr, col, val, adj.shape[0], adj.shape[1])
        if len(_dense_cache) > 64:
            _dense_cache.clear()
        try:
            _dense_cache[key] = (weakref.ref(adj), adj._version, g)
        except TypeError:
            pass
        return g

    @staticmethod
    def from_coo(edges: Tensor, w: Optional[Tensor], n_rows: int, n_cols: Optional[int] = None,
                 deterministic: bool = False) -> "CSRGraph":
        """edges[E,2] int64 ([:,0]=src, [:,1]=dst; RAGraph_edge/modules/RAGraph.py:22-24) -> CSR by dst.
        deterministic=True orders each row by original edge position (stable sort) so the fp32 sum order
        is reproducible; the default atomic-cursor build is faster and, like the reference's scatter_add_,
        fixes no order."""
        n_cols = n_rows if n_cols is None else n_cols
        if deterministic:
            dst_sorted, perm = torch.sort(edges[:, 1], stable=True)
            counts = torch.bincount(dst_sorted, minlength=n_rows)
            rowptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=edges.device)
            torch.cumsum(counts, 0, out=rowptr[1:])
            col = edges[:, 0][perm].to(torch.int32)
            val = None if w is None else w[perm].contiguous()
            return CSRGraph(rowptr, col, val, n_rows, n_cols)
        rowptr, col, val = ops.csr_from_coo(edges, w, n_rows)
        return CSRGraph(rowptr, col, val, n_rows, n_cols)

    @staticmethod
    def from_torch_sparse(adj: Tensor) -> "CSRGraph":
        """torch sparse COO/CSR tensor (the `sparse=True` branch of layers/gcn.py:33-34)."""
        if adj.layout == torch.sparse_csr:
            return CSRGraph(adj.crow_indices().to(torch.int64), adj.col_indices().to(torch.int32),
                            adj.values().to(torch.float32), adj.shape[0], adj.shape[1])
        adj = adj.coalesce()
        ind = adj.indices()
        edges = torch.stack([ind[1], ind[0]], dim=1).contiguous()      # row = dst, col = src
        return CSRGraph.from_coo(edges, adj.values().to(torch.float32), adj.shape[0], adj.shape[1],
                                 deterministic=True)

    def spmm(self, x: Tensor, epilogue: int = 0, **kw) -> Tensor:
        """epilogue(A @ x): one fused launch; differentiable (dX = A^T dY on the cached transpose) when an input
        requires grad -- see autograd.py."""
        from .autograd import spmm_epilogue
        return spmm_epilogue(self, x, epilogue, **kw)

    def to_dense(self) -> Tensor:
        """dense [n_rows, n_cols] float32 matrix (duplicate entries add up) -- for the consumers that are dense by nature
        (Augmentation's random edge rewrite, PositionAwareEncoder's all-pairs distances, sub-adjacency extraction on toy
        graphs of a few dozen nodes)."""
        val = self.val if self.val is not None else torch.ones(self.nnz, dtype=torch.float32, device=self.col.device)
        out = torch.zeros((self.n_rows, self.n_cols), dtype=torch.float32, device=self.col.device)
        out.index_put_((self.row_ids(), self.col.to(torch.int64)), val, accumulate=True)
        return out

    def row_ids(self) -> Tensor:
        """int64 [nnz]: the row of every stored entry."""
        counts = self.rowptr[1:] - self.rowptr[:-1]
        return torch.repeat_interleave(torch.arange(self.n_rows, device=self.col.device), counts)

    def transpose(self) -> "CSRGraph":
        """CSR of A^T (cached).  Entries of a transposed row are ordered by original row (stable sort), so the
        backward SpMM `dX = A^T dY` sums in a fixed order."""
        t = getattr(self, "_transposed", None)
        if t is None:
            edges_t = torch.stack([self.row_ids(), self.col.to(torch.int64)], dim=1)     # [:,0] = new col, [:,1] = new row
            t = CSRGraph.from_coo(edges_t, self.val, self.n_cols, self.n_rows, deterministic=True)
            t._transposed = self
            self._transposed = t
        return t

    def row_normalized(self) -> "CSRGraph":
        """val[i,j] / sum_j val[i,j] (Propagation.py:15-16) as its own CSR handle (cached); 0/0 rows give NaN as in
        the reference's `adj / degree`."""
        r = getattr(self, "_row_normalized", None)
        if r is None:
            val = self.val if self.val is not None else torch.ones(self.nnz, dtype=torch.float32, device=self.col.device)
            rows = self.row_ids()
            deg = torch.zeros(self.n_rows, dtype=torch.float32, device=val.device).index_add_(0, rows, val)
            r = CSRGraph(self.rowptr, self.col, val / deg[rows], self.n_rows, self.n_cols)
            self._row_normalized = r
        return r


_dense_cache: dict = {}


def as_dense(adj) -> Tensor:
    """dense 2-D view of any accepted adjacency (CSRGraph, torch sparse, dense [n,n] or [1,n,n])"""
    if isinstance(adj, CSRGraph):
        return adj.to_dense()
    if isinstance(adj, Tensor):
        if adj.layout != torch.strided:
            return adj.to_dense()
        return adj[0] if adj.dim() == 3 and adj.shape[0] == 1 else adj
    raise TypeError(f"adjacency must be a dense/sparse tensor or CSRGraph, got {type(adj)}")


def as_csr(adj) -> CSRGraph:
    if isinstance(adj, CSRGraph):
        return adj
    if isinstance(adj, Tensor):
        if adj.layout != torch.strided:
            return CSRGraph.from_torch_sparse(adj)
        return CSRGraph.from_dense(adj)
    raise TypeError(f"adjacency must be a dense/sparse tensor or CSRGraph, got {type(adj)}")
