"""Key-row sharded retrieval across the GPUs of one box (new functionality).

One process per GPU (torchrun).  Rank r holds library rows [lo_r, hi_r); the query batch is replicated.
Each rank runs the fused similarity + top-k kernel over its rows (indices already global through
idx_offset), the per-rank candidates travel in ONE all-gather (NCCL over NVLink/NVSwitch; Q*k*12 bytes
per rank), and every rank merges R*k -> k with the deterministic order (score desc, index asc), so all
ranks hold identical results.  Values/labels: owners gather the winning rows they hold into a
[Q,k,d] buffer, one all-gather, and a select by owner -- bit exact.

On NCCL/CUDA groups ``retrieve()`` fuses everything after the local top-k into ONE kernel over NVLink peer
memory (csrc/exchange.cu: candidates pushed into every peer's block, flag wait, merge, owners store their
winning rows straight into every peer's result block, flag wait) -- no NCCL call, no [R,Q,k,d] staging.  The
all-gather formulation above stays as the generic path (``topk()`` / ``gather()``, and gloo).

The compute callables are injectable so the host logic (partition, offsets, exchange, merge order) is
testable with the gloo backend on CPU; the defaults are the CUDA ops and there is no CPU fallback in
the product path.
"""
from __future__ import annotations

import os
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
    return store.topk_local(q, k, store.shard_lo)


def _default_merge(scores, idx, k):
    from . import ops
    return ops.topk_merge(scores, idx, k)


def _default_gather_owned(table_local, idx, lo, n_global, out):
    from . import ops
    ops.gather_rows_owned(table_local, idx, lo, n_global, out)


class _PeerWorkspace:
    """This rank's block of a symmetric allocation + the peer pointer table (torch symmetric memory)."""

    def __init__(self, group, device, q_max: int, k_max: int, world: int, rank: int, rba: int, rbb: int):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _lib as L
        offs = (C.c_size_t * 8)()
        L.check(L.load().rag_xchg_layout(q_max, k_max, world, rba, rbb, offs), "xchg_layout")
        (self.total, self.off_flags, self.off_a, self.par_a, self.off_b, self.par_b, _, _) = [int(x) for x in offs]
        self.key = (q_max, k_max, rba, rbb)
        self.buf = symm.empty(self.total, dtype=torch.uint8, device=device)
        self.hdl = symm.rendezvous(self.buf, group.group_name if hasattr(group, "group_name") else group)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group=group)                 # nobody signals before every block is zeroed
        self.peers_dev = int(self.hdl.buffer_ptrs_dev)
        self.step = 0

    def result(self, which: str, q: int, k: int, row_shape, dtype) -> Tensor:
        off, par = (self.off_a, self.par_a) if which == "a" else (self.off_b, self.par_b)
        n = q * k
        rb = dtype.itemsize
        for s_ in row_shape:
            rb *= s_
        o = off + (self.step & 1) * par
        return self.buf[o:o + n * rb].view(dtype).view((q, k) + tuple(row_shape))


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
        self._pws: Optional[_PeerWorkspace] = None
        self._p2p_disabled = os.environ.get("RAG_P2P", "1") == "0" or local_topk is not _default_local_topk
        self.last_path = "single"

    # ---- fused finish over NVLink peer memory ----------------------------------------------------
    def _p2p_usable(self, q: Tensor) -> bool:
        if self._p2p_disabled or self.world == 1 or not q.is_cuda:
            return False
        return dist.get_backend(self.group) == "nccl"

    def _peer_ws(self, Q: int, k: int, rba: int, rbb: int, device) -> Optional[_PeerWorkspace]:
        w = self._pws
        if w is not None and w.key[0] >= Q and w.key[1] >= k and w.key[2:] == (rba, rbb):
            return w
        grp = self.group if self.group is not None else dist.group.WORLD
        try:
            import torch.distributed._symmetric_memory  # noqa: F401  (absent from older torch builds)
        except ImportError as e:
            return self._no_p2p(e)
        try:
            # every rank takes this branch for the same call (same Q, k, row sizes), so the rendezvous matches
            self._pws = _PeerWorkspace(grp, device, max(Q, w.key[0] if w else 0), max(k, w.key[1] if w else 0),
                                       self.world, self.rank, rba, rbb)
        except (RuntimeError, NotImplementedError) as e:
            # what symmetric-memory allocation / rendezvous raises on a box without peer access (no NVLink P2P, MIG, ...):
            # fall back to the NCCL formulation, loudly.  Anything else (a bug in this library, a bad argument) propagates.
            from ._lib import RagError
            if isinstance(e, RagError):
                raise
            return self._no_p2p(e)
        return self._pws

    def _no_p2p(self, e) -> None:
        import warnings
        warnings.warn(f"ragraph_b200: peer-memory exchange unavailable ({e!r}); using the NCCL all-gather path")
        self._p2p_disabled = True
        self._pws = None
        return None

    def _finish_p2p(self, s: Tensor, i: Tensor, k: int, copy: bool):
        import ctypes as C
        from . import _lib as L
        store = self.store
        va, lb = store.resource_values, store.resource_labels
        rba = va.shape[1:].numel() * va.element_size()
        rbb = lb.shape[1:].numel() * lb.element_size()
        Q = s.shape[0]
        ws = self._peer_ws(Q, k, rba, rbb, s.device)
        if ws is None:
            return None
        ws.step += 1
        out_s = torch.empty((Q, k), dtype=torch.float32, device=s.device)
        out_i = torch.empty((Q, k), dtype=torch.int64, device=s.device)
        s, i = s.contiguous(), i.contiguous()
        with torch.cuda.device(s.device):
            L.check(L.load().rag_sharded_finish(
                C.c_void_p(s.data_ptr()), C.c_void_p(i.data_ptr()), Q, k, self.world, self.rank,
                C.c_void_p(ws.peers_dev), ws.key[0], ws.key[1], C.c_void_p(va.data_ptr()), rba,
                C.c_void_p(lb.data_ptr()), rbb, self.lo, self.hi, ws.step, C.c_void_p(out_s.data_ptr()),
                C.c_void_p(out_i.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "sharded_finish")
        emb = ws.result("a", Q, k, va.shape[1:], va.dtype)
        lab = ws.result("b", Q, k, lb.shape[1:], lb.dtype)
        if copy:
            emb, lab = emb.clone(), lab.clone()
        return emb, lab, out_s, out_i

    def _local(self, q: Tensor, k: int) -> Tuple[Tensor, Tensor]:
        k_local = min(k, self.hi - self.lo)
        s, i = self._local_topk(self.store, q, k_local)
        if k_local < k:                     # tiny shard: pad with never-winning candidates
            pad = k - k_local
            s = torch.cat([s, s.new_full((s.shape[0], pad), -torch.finfo(torch.float32).max)], 1)
            i = torch.cat([i, i.new_full((i.shape[0], pad), -1)], 1)
        return s, i

    def _exchange_nccl(self, s: Tensor, i: Tensor, k: int) -> Tuple[Tensor, Tensor]:
        # concatenated layout ([R*Q, k]) is the form both NCCL and gloo accept; viewed as [R, Q, k]
        all_s = torch.empty((self.world * s.shape[0], s.shape[1]), dtype=s.dtype, device=s.device)
        all_i = torch.empty((self.world * i.shape[0], i.shape[1]), dtype=i.dtype, device=i.device)
        dist.all_gather_into_tensor(all_s, s.contiguous(), group=self.group)
        dist.all_gather_into_tensor(all_i, i.contiguous(), group=self.group)
        return self._merge(all_s.view(self.world, *s.shape), all_i.view(self.world, *i.shape), k)

    def topk(self, q: Tensor, k: int) -> Tuple[Tensor, Tensor]:
        s, i = self._local(q, k)
        if self.world == 1:
            return s, i
        return self._exchange_nccl(s, i, k)

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

    def retrieve(self, q: Tensor, k: Optional[int] = None, copy: bool = True, events=None, wait_event=None):
        """(rag_embeddings[Q,k,d], rag_labels[Q,k,C], scores, idx) for the sharded library.
        copy=False returns views of the peer-memory result block (valid until the next-but-one call).
        events = (start, end) CUDA events recorded around the LOCAL fused top-k (bench roofline).
        wait_event: a CUDA event the stream waits for between the local top-k and the finish kernel.  A caller that reads
        the copy=False views of call i on ANOTHER stream (an overlapped device-to-host copy) must pass that read's event to
        call i+1: finishing call i+1 is what lets the peers start call i+2, whose rows land in the block call i used."""
        if q.dim() == 1:
            q = q.unsqueeze(0)
        k = self.store.retrieve_num if k is None else k
        if events:
            events[0].record()
        s, i = self._local(q, k)
        if events:
            events[1].record()
        if wait_event is not None:
            torch.cuda.current_stream().wait_event(wait_event)
        if self._p2p_usable(q):
            out = self._finish_p2p(s, i, k, copy)
            if out is not None:
                self.last_path = "p2p"
                return out
            # the rendezvous failed on every rank alike: finish this call with the NCCL formulation
        scores, idx = (s, i) if self.world == 1 else self._exchange_nccl(s, i, k)
        self.last_path = "nccl" if self.world > 1 else "single"
        return (self.gather(self.store.resource_values, idx), self.gather(self.store.resource_labels, idx),
                scores, idx)
