"""Key-row sharded retrieval across the GPUs of one box (new functionality).

One process per GPU (torchrun).  Rank r holds library rows [lo_r, hi_r); the query batch is replicated.
Each rank runs the fused similarity + top-k kernel over its rows (indices already global through
idx_offset), the per-rank candidates travel in ONE all-gather (NCCL over NVLink/NVSwitch; Q*k*12 bytes
per rank), and every rank merges R*k -> k with the deterministic order (score desc, index asc), so all
ranks hold identical results.  Values/labels: owners gather the winning rows they hold into a
[Q,k,d] buffer, one all-gather, and a select by owner -- bit exact.

The compute callables are injectable so the host logic (partition, offsets, exchange, merge order) is
testable with the gloo backend on CPU; the defaults are the CUDA ops and there is no CPU fallback in
the product path.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced row partition: the first n_rows % world ranks hold one extra row."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owner_of(idx: Tensor, n_rows: int, world: int) -> Tensor:
    """rank that owns each global row index under shard_bounds."""
    base, rem = divmod(n_rows, world)
    cut = rem * (base + 1)
    if base == 0:
        return idx.clone()
    return torch.where(idx < cut, idx // (base + 1), rem + (idx - cut) // base)


def _default_local_topk(store, q, k):
    from . import ops
    mode = store._pick_mode(q.shape[0], k)
    store._refresh_derived(mode != 0)
    return ops.cosine_topk(q, store.resource_keys, k, store._inv_norm[:len(store)],
                           store._keys_bf16 if mode != 0 else None, mode, 0, store.shard_lo)


def _default_merge(scores, idx, k):
    from . import ops
    return ops.topk_merge(scores, idx, k)


def _default_gather_owned(table_local, idx, lo, n_global, out):
    from . import ops
    ops.gather_rows_owned(table_local, idx, lo, n_global, out)


class ShardedRetriever:
    """Wraps a per-rank store (ToyGraphBase holding this rank's rows, with ``shard_lo`` set) and a process group."""

    def __init__(self, store, n_global: int, group: Optional[dist.ProcessGroup] = None,
                 local_topk: Callable = _default_local_topk, merge: Callable = _default_merge,
                 gather_owned: Callable = _default_gather_owned):
        self.store, self.n_global, self.group = store, n_global, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lo, self.hi = shard_bounds(n_global, self.world, self.rank)
        store.shard_lo = self.lo
        self._local_topk, self._merge, self._gather_owned = local_topk, merge, gather_owned

    def topk(self, q: Tensor, k: int) -> Tuple[Tensor, Tensor]:
        k_local = min(k, self.hi - self.lo)
        s, i = self._local_topk(self.store, q, k_local)
        if k_local < k:                     # tiny shard: pad with never-winning candidates
            pad = k - k_local
            s = torch.cat([s, s.new_full((s.shape[0], pad), -torch.finfo(torch.float32).max)], 1)
            i = torch.cat([i, i.new_full((i.shape[0], pad), -1)], 1)
        if self.world == 1:
            return s, i
        # concatenated layout ([R*Q, k]) is the form both NCCL and gloo accept; viewed as [R, Q, k]
        all_s = torch.empty((self.world * s.shape[0], s.shape[1]), dtype=s.dtype, device=s.device)
        all_i = torch.empty((self.world * i.shape[0], i.shape[1]), dtype=i.dtype, device=i.device)
        dist.all_gather_into_tensor(all_s, s.contiguous(), group=self.group)
        dist.all_gather_into_tensor(all_i, i.contiguous(), group=self.group)
        return self._merge(all_s.view(self.world, *s.shape), all_i.view(self.world, *i.shape), k)

    def gather(self, table_local: Tensor, idx: Tensor) -> Tensor:
        """rows of the GLOBAL table addressed by idx [Q,k]; table_local = this rank's rows."""
        out = torch.zeros(tuple(idx.shape) + tuple(table_local.shape[1:]), dtype=table_local.dtype,
                          device=table_local.device)
        self._gather_owned(table_local, idx.contiguous(), self.lo, self.n_global, out)
        if self.world == 1:
            return out
        allr = torch.empty((self.world * out.shape[0],) + tuple(out.shape[1:]), dtype=out.dtype, device=out.device)
        dist.all_gather_into_tensor(allr, out, group=self.group)
        allr = allr.view((self.world,) + tuple(out.shape))
        own = owner_of(idx, self.n_global, self.world)                      # [Q,k]
        sel = own.reshape((1,) + tuple(idx.shape) + (1,) * (out.dim() - idx.dim()))
        return torch.gather(allr, 0, sel.expand((1,) + tuple(out.shape))).squeeze(0)

    def retrieve(self, q: Tensor, k: Optional[int] = None):
        """(rag_embeddings[Q,k,d], rag_labels[Q,k,C], scores, idx) for the sharded library."""
        if q.dim() == 1:
            q = q.unsqueeze(0)
        k = self.store.retrieve_num if k is None else k
        scores, idx = self.topk(q, k)
        return (self.gather(self.store.resource_values, idx), self.gather(self.store.resource_labels, idx),
                scores, idx)
