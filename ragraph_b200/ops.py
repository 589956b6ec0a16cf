"""torch custom ops ``torch.ops.ragraph.*`` -- a thin layer over the C ABI (include/ragraph_b200.h).

PyTorch is plumbing here: it owns device memory and the current stream; every op body is one or
two calls into libragraph_b200.so.  CUDA tensors only -- there is no CPU kernel and no PyTorch
fallback; a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib as L


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _need_cuda(*ts: Optional[Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("ragraph_b200 ops run on CUDA tensors only (no CPU fallback); got a "
                               f"{t.device} tensor")


def _f32c(t: Tensor, name: str) -> Tensor:
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name}: expected float32, got {t.dtype}")
    return t.contiguous()


def _workspace(nbytes: int, device) -> Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def direct(op):
    """The Python body of a ``torch.ops.ragraph`` op without the torch.library dispatcher in front of it.

    ``torch.library.custom_op`` costs ~15 us of host time per call (measured); a retrieve is 3-4 ops and the edge
    variant's forward runs ~120 of them, so the eager, forward-only host loops (ToyGraphBase.retrieve, the no-grad SpMM
    path, the sharded retriever) call the body directly -- same checks, same C-ABI call, same results.  The registered
    ops stay the public surface (``torch.ops.ragraph.*``, fake kernels for tracing)."""
    return getattr(op, "_init_fn", op)


# ----------------------------------------------------------------------------- norms / shadow
@torch.library.custom_op("ragraph::row_inv_norm", mutates_args=())
def row_inv_norm(x: Tensor, eps: float = 1e-12) -> Tensor:
    _need_cuda(x)
    x = _f32c(x, "row_inv_norm")
    out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.load().rag_row_inv_norm_f32(_p(x), x.shape[0], x.shape[1], eps, _p(out), _stream()), "row_inv_norm")
    return out


@row_inv_norm.register_fake
def _(x, eps=1e-12):
    return x.new_empty(x.shape[0])


@torch.library.custom_op("ragraph::rows_normalize", mutates_args=())
def rows_normalize(x: Tensor, eps: float = 1e-12) -> Tensor:
    """F.normalize(x, p=2, dim=-1) for a 2-D fp32 tensor."""
    _need_cuda(x)
    x = _f32c(x, "rows_normalize")
    if x.dim() != 2:
        raise RuntimeError("rows_normalize: 2-D tensor expected")
    out = torch.empty_like(x)
    if x.numel():
        with torch.cuda.device(x.device):
            L.check(L.load().rag_rows_normalize_f32(_p(x), x.shape[0], x.shape[1], eps, _p(out), _stream()), "rows_normalize")
    return out


@rows_normalize.register_fake
def _(x, eps=1e-12):
    return torch.empty_like(x)


@torch.library.custom_op("ragraph::rows_to_bf16", mutates_args=())
def rows_to_bf16(x: Tensor, normalize: bool = True, eps: float = 1e-12) -> Tensor:
    """bf16 (optionally L2-normalised) shadow [rows, round_up(d, 64)] for the tensor-core filter."""
    _need_cuda(x)
    x = _f32c(x, "rows_to_bf16")
    d_pad = round_up(x.shape[1], 64)
    out = torch.empty((x.shape[0], d_pad), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.load().rag_rows_to_bf16(_p(x), x.shape[0], x.shape[1], int(normalize), eps, _p(out), d_pad,
                                          _stream()), "rows_to_bf16")
    return out


@rows_to_bf16.register_fake
def _(x, normalize=True, eps=1e-12):
    return x.new_empty((x.shape[0], round_up(x.shape[1], 64)), dtype=torch.bfloat16)


def rows_to_shadow16(x: Tensor, fmt: int = L.FMT_F16, normalize: bool = True, eps: float = 1e-12,
                     err_max: Optional[Tensor] = None, want_row_err: bool = False):
    """16-bit shadow [rows, round_up(d, 64)] (fmt FMT_F16 -> float16, FMT_BF16 -> bfloat16) plus the rounding-error norms of
    the exactness certificate: ``err_max`` (float32 [1], zero-initialised by the caller, max-accumulated) and, if
    ``want_row_err``, the per-row norms.  Returns (shadow, row_err or None)."""
    _need_cuda(x, err_max)
    x = _f32c(x, "rows_to_shadow16")
    d_pad = round_up(x.shape[1], 64)
    out = torch.empty((x.shape[0], d_pad), dtype=torch.float16 if fmt == L.FMT_F16 else torch.bfloat16, device=x.device)
    row_err = torch.empty(x.shape[0], dtype=torch.float32, device=x.device) if want_row_err else None
    if err_max is not None and (err_max.dtype != torch.float32 or err_max.numel() != 1):
        raise RuntimeError("rows_to_shadow16: err_max must be a float32 tensor with one element")
    with torch.cuda.device(x.device):
        L.check(L.load().rag_rows_to_shadow16(_p(x), x.shape[0], x.shape[1], fmt, int(normalize), eps, _p(out), d_pad,
                                              _p(row_err), _p(err_max), _stream()), "rows_to_shadow16")
    return out, row_err


def tf32_shadow_dpad(d: int) -> int:
    """columns of the tf32 key shadow for embedding dim d (32, 64 or 128; 0 = d not covered by RAG_SIM_TF32)"""
    return int(L.load().rag_tf32_shadow_dpad(d))


@torch.library.custom_op("ragraph::rows_to_tf32", mutates_args=())
def rows_to_tf32(x: Tensor, normalize: bool = True, eps: float = 1e-12) -> Tensor:
    """tf32-rounded (optionally L2-normalised) fp32 shadow [rows, tf32_shadow_dpad(d)] for RAG_SIM_TF32."""
    _need_cuda(x)
    x = _f32c(x, "rows_to_tf32")
    d_pad = tf32_shadow_dpad(x.shape[1])
    if d_pad == 0:
        raise RuntimeError(f"rows_to_tf32: the tf32 mode covers d <= 128, got d={x.shape[1]}")
    out = torch.empty((x.shape[0], d_pad), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.load().rag_rows_to_tf32(_p(x), x.shape[0], x.shape[1], int(normalize), eps, _p(out), d_pad,
                                          _stream()), "rows_to_tf32")
    return out


@rows_to_tf32.register_fake
def _(x, normalize=True, eps=1e-12):
    d = x.shape[1]
    return x.new_empty((x.shape[0], 32 if d <= 32 else (64 if d <= 64 else 128)))


# ----------------------------------------------------------------------------- similarity
@torch.library.custom_op("ragraph::cosine_similarity", mutates_args=())
def cosine_similarity(q: Tensor, keys: Tensor, flags: int = 0) -> Tensor:
    _need_cuda(q, keys)
    q, keys = _f32c(q, "cosine_similarity"), _f32c(keys, "cosine_similarity")
    if q.shape[1] != keys.shape[1]:
        raise RuntimeError(f"cosine_similarity: dim mismatch {tuple(q.shape)} vs {tuple(keys.shape)}")
    Q, N, d = q.shape[0], keys.shape[0], q.shape[1]
    out = torch.empty((Q, N), dtype=torch.float32, device=q.device)
    lib = L.load()
    ws = _workspace(lib.rag_cosine_similarity_workspace(Q, N), q.device)
    with torch.cuda.device(q.device):
        L.check(lib.rag_cosine_similarity_f32(_p(q), Q, _p(keys), N, d, flags, _p(out), _p(ws), ws.numel(),
                                              _stream()), "cosine_similarity")
    return out


@cosine_similarity.register_fake
def _(q, keys, flags=0):
    return q.new_empty((q.shape[0], keys.shape[0]))


_SHADOW_DTYPE = {L.SIM_BF16: torch.bfloat16, L.SIM_BF16_REFINE: torch.bfloat16, L.SIM_F16: torch.float16,
                 L.SIM_F16_REFINE: torch.float16}


def _cosine_topk_impl(q, keys, k, key_inv_norm, keys_shadow, mode, flags, idx_offset, shadow_err, want_stats):
    _need_cuda(q, keys, key_inv_norm, keys_shadow, shadow_err)
    q, keys = _f32c(q, "cosine_topk"), _f32c(keys, "cosine_topk")
    if q.dim() != 2 or keys.dim() != 2 or q.shape[1] != keys.shape[1]:
        raise RuntimeError(f"cosine_topk: shapes {tuple(q.shape)} vs {tuple(keys.shape)}")
    Q, N, d = q.shape[0], keys.shape[0], q.shape[1]
    if key_inv_norm is not None:
        key_inv_norm = _f32c(key_inv_norm, "key_inv_norm")
        if key_inv_norm.numel() != N:
            raise RuntimeError("cosine_topk: key_inv_norm must have N entries")
    if keys_shadow is not None and mode == L.SIM_TF32:
        if keys_shadow.dtype != torch.float32 or tuple(keys_shadow.shape) != (N, tf32_shadow_dpad(d)) \
                or not keys_shadow.is_contiguous():
            raise RuntimeError("cosine_topk: SIM_TF32 needs the contiguous [N, tf32_shadow_dpad(d)] fp32 shadow (rows_to_tf32)")
    elif keys_shadow is not None:
        want = _SHADOW_DTYPE.get(mode, keys_shadow.dtype)
        if keys_shadow.dtype != want or tuple(keys_shadow.shape) != (N, round_up(d, 64)) or not keys_shadow.is_contiguous():
            raise RuntimeError(f"cosine_topk: mode {mode} needs the contiguous [N, round_up(d,64)] {want} shadow "
                               f"(rows_to_shadow16), got {keys_shadow.dtype} {tuple(keys_shadow.shape)}")
    if shadow_err is not None and (shadow_err.dtype != torch.float32 or shadow_err.numel() != 1):
        raise RuntimeError("cosine_topk: shadow_err must be the float32 [1] err_max of rows_to_shadow16")
    scores = torch.empty((Q, k), dtype=torch.float32, device=q.device)
    idx = torch.empty((Q, k), dtype=torch.int64, device=q.device)
    lib = L.load()
    ws = _workspace(lib.rag_cosine_topk_workspace(Q, N, d, k, mode), q.device)
    with torch.cuda.device(q.device):
        L.check(lib.rag_cosine_topk_f32(_p(q), Q, _p(keys), _p(key_inv_norm), _p(keys_shadow), _p(shadow_err), N, d, k,
                                        mode, flags, idx_offset, _p(scores), _p(idx), _p(ws), ws.numel(), _stream()),
                "cosine_topk")
    if not want_stats:
        return scores, idx
    offs = (C.c_size_t * 3)()
    L.check(lib.rag_cosine_topk_stat_offsets(Q, N, d, k, mode, offs), "cosine_topk_stat_offsets")
    if offs[0] == 0 and offs[1] == 0:
        stats = torch.zeros(5, dtype=torch.int32, device=q.device)
    else:
        stats = ws[offs[0]:offs[0] + 20].view(torch.int32)         # the five counters are consecutive; view keeps ws alive
    return scores, idx, stats


@torch.library.custom_op("ragraph::cosine_topk", mutates_args=())
def cosine_topk(q: Tensor, keys: Tensor, k: int, key_inv_norm: Optional[Tensor] = None,
                keys_bf16: Optional[Tensor] = None, mode: int = 0, flags: int = 0,
                idx_offset: int = 0, shadow_err: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """Fused similarity + top-k.  Returns (scores[Q,k] f32 desc, idx[Q,k] int64).  ``keys_bf16`` is the key shadow of
    the tensor-core modes: rows_to_shadow16(keys, FMT_F16) for SIM_F16 / SIM_F16_REFINE, rows_to_bf16(keys) for SIM_BF16 /
    SIM_BF16_REFINE, rows_to_tf32(keys) for SIM_TF32; ``shadow_err`` = that shadow's err_max (tightens the certificate)."""
    return _cosine_topk_impl(q, keys, k, key_inv_norm, keys_bf16, mode, flags, idx_offset, shadow_err, False)


@cosine_topk.register_fake
def _(q, keys, k, key_inv_norm=None, keys_bf16=None, mode=0, flags=0, idx_offset=0, shadow_err=None):
    return q.new_empty((q.shape[0], k)), q.new_empty((q.shape[0], k), dtype=torch.int64)


def cosine_topk_with_stats(q: Tensor, keys: Tensor, k: int, key_inv_norm: Optional[Tensor] = None,
                           keys_shadow: Optional[Tensor] = None, mode: int = 0, flags: int = 0, idx_offset: int = 0,
                           shadow_err: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """cosine_topk plus an int32 [5] device tensor {rows that took the second tensor-core pass, rows that fell back to the
    fp32 kernel, rows whose certificate would fail under an 8x error bound, worker CTAs counted by the cross-split sweep
    (0 = the sweep did not run), (lane, 32-score chunk) hits the query-stationary filter queued} (zeros for modes without
    a certificate)."""
    return _cosine_topk_impl(q, keys, k, key_inv_norm, keys_shadow, mode, flags, idx_offset, shadow_err, True)


def retrieve_small_supported(Q: int, N: int, d: int, k: int) -> bool:
    return bool(L.load().rag_retrieve_small_supported(Q, N, d, k))


def retrieve_small_workspace(Q: int, N: int, d: int, k: int, device) -> Tensor:
    """zero-filled scratch for retrieve_small (reusable across calls of the same or smaller shape on one stream)"""
    return torch.zeros(max(int(L.load().rag_retrieve_small_workspace(Q, N, d, k)), 256), dtype=torch.uint8, device=device)


def retrieve_small(q: Tensor, keys: Tensor, k: int, values: Optional[Tensor], labels: Optional[Tensor], ws: Tensor,
                   key_inv_norm: Optional[Tensor] = None, flags: int = 0):
    """ToyGraphBase.retrieve for a small problem in ONE launch: returns (scores [Q,k], idx [Q,k], values[idx], labels[idx]).
    All tensors CUDA + contiguous (fp32 q / keys; any 4-byte-multiple row type for values / labels); ``ws`` from
    retrieve_small_workspace.  Host-lean on purpose: this path exists to beat launch latency."""
    if not (q.is_cuda and keys.is_cuda and q.dtype == torch.float32 and keys.dtype == torch.float32
            and q.is_contiguous() and keys.is_contiguous() and q.dim() == 2 and keys.dim() == 2 and q.shape[1] == keys.shape[1]):
        raise RuntimeError("retrieve_small: contiguous CUDA float32 q [Q,d] and keys [N,d] expected (no CPU fallback)")
    Q, d = q.shape
    N = keys.shape[0]
    dev = q.device
    scores = torch.empty((Q, k), dtype=torch.float32, device=dev)
    idx = torch.empty((Q, k), dtype=torch.int64, device=dev)
    ov = ol = None
    vb = lb = 0
    if values is not None:
        if not values.is_contiguous():
            raise RuntimeError("retrieve_small: values must be contiguous")
        ov = torch.empty((Q, k) + tuple(values.shape[1:]), dtype=values.dtype, device=dev)
        vb = _row_bytes(values)
    if labels is not None:
        if not labels.is_contiguous():
            raise RuntimeError("retrieve_small: labels must be contiguous")
        ol = torch.empty((Q, k) + tuple(labels.shape[1:]), dtype=labels.dtype, device=dev)
        lb = _row_bytes(labels)
    with torch.cuda.device(dev):
        L.check(L.load().rag_retrieve_small_f32(q.data_ptr(), Q, keys.data_ptr(), _p(key_inv_norm), N, d, k, flags, _p(values), vb,
                                                _p(labels), lb, scores.data_ptr(), idx.data_ptr(), _p(ov), _p(ol), ws.data_ptr(),
                                                ws.numel(), _stream()), "retrieve_small")
    return scores, idx, ov, ol


@torch.library.custom_op("ragraph::topk_masked", mutates_args=())
def topk_masked(q: Tensor, keys: Tensor, k: int, mask_rowptr: Tensor, mask_col: Tensor, flags: int = 0,
                key_inv_norm: Optional[Tensor] = None, idx_offset: int = 0, mode: int = 0,
                keys_shadow: Optional[Tensor] = None, shadow_err: Optional[Tensor] = None,
                key_scale: float = 0.0) -> Tuple[Tensor, Tensor]:
    """Fused similarity + top-k where query row r never returns the key indices mask_col[mask_rowptr[r]:mask_rowptr[r+1]]
    (int64 CSR of exclusions).  flags=SIM_DOT: plain dot product (edge evaluation ranking).
    mode = SIM_F16_REFINE / SIM_BF16_REFINE runs it on the tensor cores (exact): ``keys_shadow`` / ``shadow_err`` from
    rows_to_shadow16 -- of the normalised keys (cosine; ``key_inv_norm`` required) or, with SIM_DOT, of ``keys * key_scale``
    un-normalised with ``key_scale`` = 1 / (largest key norm)."""
    _need_cuda(q, keys, mask_rowptr, mask_col, key_inv_norm, keys_shadow, shadow_err)
    q, keys = _f32c(q, "topk_masked"), _f32c(keys, "topk_masked")
    if q.dim() != 2 or keys.dim() != 2 or q.shape[1] != keys.shape[1]:
        raise RuntimeError(f"topk_masked: shapes {tuple(q.shape)} vs {tuple(keys.shape)}")
    Q, N, d = q.shape[0], keys.shape[0], q.shape[1]
    if mask_rowptr.dtype != torch.int64 or mask_col.dtype != torch.int64 or mask_rowptr.numel() != Q + 1:
        raise RuntimeError("topk_masked: mask_rowptr int64[Q+1] and mask_col int64[nnz] expected")
    mask_rowptr, mask_col = mask_rowptr.contiguous(), mask_col.contiguous()
    if key_inv_norm is not None:
        key_inv_norm = _f32c(key_inv_norm, "key_inv_norm")
    scores = torch.empty((Q, k), dtype=torch.float32, device=q.device)
    idx = torch.empty((Q, k), dtype=torch.int64, device=q.device)
    lib = L.load()
    if mode == L.SIM_FP32:
        ws = _workspace(lib.rag_cosine_topk_workspace(Q, N, d, k, L.SIM_FP32), q.device)
        with torch.cuda.device(q.device):
            L.check(lib.rag_topk_masked_f32(_p(q), Q, _p(keys), _p(key_inv_norm), N, d, k, flags, _p(mask_rowptr), _p(mask_col),
                                            idx_offset, _p(scores), _p(idx), _p(ws), ws.numel(), _stream()), "topk_masked")
        return scores, idx
    want = torch.float16 if mode == L.SIM_F16_REFINE else torch.bfloat16
    if keys_shadow is None or keys_shadow.dtype != want or tuple(keys_shadow.shape) != (N, round_up(d, 64)) \
            or not keys_shadow.is_contiguous():
        raise RuntimeError(f"topk_masked: mode {mode} needs the contiguous [N, round_up(d,64)] {want} shadow (rows_to_shadow16)")
    ws = _workspace(lib.rag_cosine_topk_workspace(Q, N, d, k, mode), q.device)
    with torch.cuda.device(q.device):
        L.check(lib.rag_topk_masked_tc_f32(_p(q), Q, _p(keys), _p(key_inv_norm), _p(keys_shadow), _p(shadow_err), N, d, k, mode,
                                           flags, float(key_scale), _p(mask_rowptr), _p(mask_col), idx_offset, _p(scores),
                                           _p(idx), _p(ws), ws.numel(), _stream()), "topk_masked_tc")
    return scores, idx


@topk_masked.register_fake
def _(q, keys, k, mask_rowptr, mask_col, flags=0, key_inv_norm=None, idx_offset=0, mode=0, keys_shadow=None, shadow_err=None,
      key_scale=0.0):
    return q.new_empty((q.shape[0], k)), q.new_empty((q.shape[0], k), dtype=torch.int64)


@torch.library.custom_op("ragraph::cosine2_topk", mutates_args=())
def cosine2_topk(qa: Tensor, ka: Tensor, w_a: float, qb: Tensor, kb: Tensor, w_b: float,
                 k: int) -> Tuple[Tensor, Tensor]:
    """top-k of w_a*cos(qa,ka) + w_b*cos(qb,kb) (node_fewshot two-metric retrieval)."""
    _need_cuda(qa, ka, qb, kb)
    qa, ka, qb, kb = (_f32c(t, "cosine2_topk") for t in (qa, ka, qb, kb))
    Q, N = qa.shape[0], ka.shape[0]
    if qb.shape[0] != Q or kb.shape[0] != N or qa.shape[1] != ka.shape[1] or qb.shape[1] != kb.shape[1]:
        raise RuntimeError("cosine2_topk: inconsistent shapes")
    scores = torch.empty((Q, k), dtype=torch.float32, device=qa.device)
    idx = torch.empty((Q, k), dtype=torch.int64, device=qa.device)
    lib = L.load()
    ws = _workspace(lib.rag_cosine2_topk_workspace(Q, N, qa.shape[1], qb.shape[1], k), qa.device)
    with torch.cuda.device(qa.device):
        L.check(lib.rag_cosine2_topk_f32(_p(qa), _p(ka), qa.shape[1], w_a, _p(qb), _p(kb), qb.shape[1], w_b, Q, N, k,
                                         _p(scores), _p(idx), _p(ws), ws.numel(), _stream()), "cosine2_topk")
    return scores, idx


@cosine2_topk.register_fake
def _(qa, ka, w_a, qb, kb, w_b, k):
    return qa.new_empty((qa.shape[0], k)), qa.new_empty((qa.shape[0], k), dtype=torch.int64)


@torch.library.custom_op("ragraph::topk_merge", mutates_args=())
def topk_merge(scores: Tensor, idx: Tensor, k_out: int) -> Tuple[Tensor, Tensor]:
    """[R,Q,k_in] per-shard candidates -> global [Q,k_out] (score desc, index asc)."""
    _need_cuda(scores, idx)
    scores = _f32c(scores, "topk_merge")
    idx = idx.contiguous()
    if idx.dtype != torch.int64 or scores.shape != idx.shape or scores.dim() != 3:
        raise RuntimeError("topk_merge: scores f32 [R,Q,k] and idx int64 [R,Q,k] expected")
    R, Q, k_in = scores.shape
    out_s = torch.empty((Q, k_out), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((Q, k_out), dtype=torch.int64, device=scores.device)
    with torch.cuda.device(scores.device):
        L.check(L.load().rag_topk_merge(_p(scores), _p(idx), R, Q, k_in, k_out, _p(out_s), _p(out_i), _stream()),
                "topk_merge")
    return out_s, out_i


@topk_merge.register_fake
def _(scores, idx, k_out):
    return scores.new_empty((scores.shape[1], k_out)), idx.new_empty((scores.shape[1], k_out))


# ----------------------------------------------------------------------------- gathers
def _row_bytes(table: Tensor) -> int:
    n = table.element_size()
    for s in table.shape[1:]:
        n *= s
    return n


@torch.library.custom_op("ragraph::gather_rows", mutates_args=())
def gather_rows(table: Tensor, idx: Tensor) -> Tensor:
    """``table[idx]`` (advanced indexing on dim 0), bit exact, any dtype."""
    _need_cuda(table, idx)
    if idx.dtype != torch.int64:
        raise RuntimeError("gather_rows: idx must be int64")
    table, idx = table.contiguous(), idx.contiguous()
    out = torch.empty(tuple(idx.shape) + tuple(table.shape[1:]), dtype=table.dtype, device=table.device)
    if out.numel():
        N = table.shape[0]
        with torch.cuda.device(table.device):
            L.check(L.load().rag_gather_rows(_p(table), N, _row_bytes(table), _p(idx), idx.numel(), 0, N, _p(out),
                                             _stream()), "gather_rows")
    return out


@gather_rows.register_fake
def _(table, idx):
    return table.new_empty(tuple(idx.shape) + tuple(table.shape[1:]))


@torch.library.custom_op("ragraph::gather_rows_owned", mutates_args=("out",))
def gather_rows_owned(table_local: Tensor, idx: Tensor, owner_lo: int, n_global: int, out: Tensor) -> None:
    """Sharded 'owners gather': fill out[m] = table_local[idx[m]-owner_lo] for the global indices
    this shard owns ([owner_lo, owner_lo+rows)); other rows of ``out`` are left untouched."""
    _need_cuda(table_local, idx, out)
    if idx.dtype != torch.int64 or not (table_local.is_contiguous() and idx.is_contiguous() and out.is_contiguous()):
        raise RuntimeError("gather_rows_owned: contiguous tensors and int64 idx expected")
    if idx.numel():
        with torch.cuda.device(out.device):
            L.check(L.load().rag_gather_rows(_p(table_local), n_global, _row_bytes(table_local), _p(idx), idx.numel(),
                                             owner_lo, owner_lo + table_local.shape[0], _p(out), _stream()),
                    "gather_rows_owned")


@torch.library.custom_op("ragraph::gather_reduce", mutates_args=())
def gather_reduce(table: Tensor, idx: Tensor, op: int = 0, blend_in: Optional[Tensor] = None,
                  blend_w: float = 0.0) -> Tensor:
    """reduce_j table[idx[q,j]] (op 0 = sum, 1 = mean), optionally (1-w)*blend_in + w*reduce."""
    _need_cuda(table, idx, blend_in)
    table = _f32c(table, "gather_reduce")
    idx = idx.contiguous()
    if idx.dtype != torch.int64 or idx.dim() != 2 or table.dim() != 2:
        raise RuntimeError("gather_reduce: table [N,d] f32 and idx [Q,k] int64 expected")
    Q, k = idx.shape
    out = torch.empty((Q, table.shape[1]), dtype=torch.float32, device=table.device)
    if blend_in is not None:
        blend_in = _f32c(blend_in, "blend_in")
        if blend_in.shape != out.shape:
            raise RuntimeError("gather_reduce: blend_in must be [Q,d]")
    if Q:
        with torch.cuda.device(table.device):
            L.check(L.load().rag_gather_reduce_f32(_p(table), table.shape[0], table.shape[1], _p(idx), Q, k, op,
                                                   _p(blend_in), blend_w, _p(out), _stream()), "gather_reduce")
    return out


@gather_reduce.register_fake
def _(table, idx, op=0, blend_in=None, blend_w=0.0):
    return table.new_empty((idx.shape[0], table.shape[1]))


# ----------------------------------------------------------------------------- CSR SpMM
@torch.library.custom_op("ragraph::csr_spmm", mutates_args=())
def csr_spmm(rowptr: Tensor, col: Tensor, val: Optional[Tensor], x: Tensor, epilogue: int = 0,
             bias: Optional[Tensor] = None, alpha: Optional[Tensor] = None, blend_in: Optional[Tensor] = None,
             blend_w: float = 0.0, accum_in: Optional[Tensor] = None) -> Tensor:
    _need_cuda(rowptr, col, val, x, bias, alpha, blend_in, accum_in)
    x = _f32c(x, "csr_spmm")
    if rowptr.dtype not in (torch.int64, torch.int32) or col.dtype != torch.int32:
        raise RuntimeError("csr_spmm: rowptr int64/int32 and col int32 expected")
    rowptr, col = rowptr.contiguous(), col.contiguous()
    n_rows, n_src, F = rowptr.numel() - 1, x.shape[0], x.shape[1]
    y = torch.empty((n_rows, F), dtype=torch.float32, device=x.device)
    val = None if val is None else _f32c(val, "val")
    bias = None if bias is None else _f32c(bias, "bias")
    alpha = None if alpha is None else _f32c(alpha, "alpha").reshape(-1)
    blend_in = None if blend_in is None else _f32c(blend_in, "blend_in")
    accum_in = None if accum_in is None else _f32c(accum_in, "accum_in")
    for t, nm in ((blend_in, "blend_in"), (accum_in, "accum_in")):
        if t is not None and t.shape != y.shape:
            raise RuntimeError(f"csr_spmm: {nm} must be [n_rows, F]")
    if n_rows:
        with torch.cuda.device(x.device):
            L.check(L.load().rag_csr_spmm_f32(_p(rowptr), int(rowptr.dtype == torch.int64), _p(col), _p(val), n_rows,
                                              n_src, col.numel(), _p(x), F, epilogue, _p(bias), _p(alpha),
                                              _p(blend_in), blend_w, _p(accum_in), _p(y), _stream()), "csr_spmm")
    return y


@csr_spmm.register_fake
def _(rowptr, col, val, x, epilogue=0, bias=None, alpha=None, blend_in=None, blend_w=0.0, accum_in=None):
    return x.new_empty((rowptr.numel() - 1, x.shape[1]))


@torch.library.custom_op("ragraph::csr_from_coo", mutates_args=())
def csr_from_coo(edges: Tensor, w: Optional[Tensor], n_rows: int) -> Tuple[Tensor, Tensor, Tensor]:
    """COO edges[E,2] int64 ([:,0]=src, [:,1]=dst) (+ weights) -> CSR grouped by dst."""
    _need_cuda(edges, w)
    if edges.dtype != torch.int64 or edges.dim() != 2 or edges.shape[1] != 2:
        raise RuntimeError("csr_from_coo: edges must be int64 [E,2]")
    edges = edges.contiguous()
    w = None if w is None else _f32c(w, "w")
    E, dev = edges.shape[0], edges.device
    lib = L.load()
    counts = torch.empty(n_rows, dtype=torch.int32, device=dev)
    rowptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=dev)
    col = torch.empty(E, dtype=torch.int32, device=dev)
    val = torch.empty(E, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.rag_coo_count_rows(_p(edges), E, n_rows, _p(counts), _stream()), "coo_count_rows")
        torch.cumsum(counts, 0, dtype=torch.int64, out=rowptr[1:])
        L.check(lib.rag_coo_fill_csr(_p(edges), _p(w), E, n_rows, _p(rowptr), _p(counts), _p(col), _p(val),
                                     _stream()), "coo_fill_csr")
    return rowptr, col, val


@csr_from_coo.register_fake
def _(edges, w, n_rows):
    E = edges.shape[0]
    return (edges.new_empty(n_rows + 1), edges.new_empty(E, dtype=torch.int32),
            edges.new_empty(E, dtype=torch.float32))


@torch.library.custom_op("ragraph::csr_from_dense", mutates_args=())
def csr_from_dense(adj: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """dense [n,m] adjacency -> CSR of its non-zeros (column order kept)."""
    _need_cuda(adj)
    adj = _f32c(adj, "csr_from_dense")
    if adj.dim() != 2:
        raise RuntimeError("csr_from_dense: 2-D adjacency expected")
    n, m, dev = adj.shape[0], adj.shape[1], adj.device
    lib = L.load()
    counts = torch.empty(n, dtype=torch.int32, device=dev)
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.rag_dense_count_rows(_p(adj), n, m, _p(counts), _stream()), "dense_count_rows")
        torch.cumsum(counts, 0, dtype=torch.int64, out=rowptr[1:])
        nnz = int(rowptr[-1].item())           # sizes col/val (one sync; CSR is built once per adjacency)
        col = torch.empty(nnz, dtype=torch.int32, device=dev)
        val = torch.empty(nnz, dtype=torch.float32, device=dev)
        L.check(lib.rag_dense_fill_csr(_p(adj), n, m, _p(rowptr), _p(col), _p(val), _stream()), "dense_fill_csr")
    return rowptr, col, val


# ----------------------------------------------------------------------------- edge time encoding (K10)
@torch.library.custom_op("ragraph::scatter_softmax", mutates_args=())
def scatter_softmax(src: Tensor, index: Tensor, dim_size: int, lo: float = 0.0, span: float = 1.0,
                    base: Optional[Tensor] = None, mix_a: float = 0.0, mix_b: float = 1.0,
                    range_dev: Optional[Tensor] = None) -> Tensor:
    """mix_a * base + mix_b * softmax over the entries sharing an index of (src - lo) / span; [E] in COO order.
    Defaults = torch_scatter.scatter_softmax(src, index, dim_size=dim_size).  range_dev = device float32 [2]
    {min, max_step}: lo / span are then read on the device (no host sync)."""
    _need_cuda(src, index, base, range_dev)
    if range_dev is not None:
        range_dev = _f32c(range_dev, "range_dev")
        if range_dev.numel() != 2:
            raise RuntimeError("scatter_softmax: range_dev must hold {min, max_step}")
    src = _f32c(src, "scatter_softmax")
    if index.dtype != torch.int64 or index.dim() != 1 or src.dim() != 1 or index.numel() != src.numel():
        raise RuntimeError("scatter_softmax: src f32 [E] and index int64 [E] expected")
    index = index.contiguous()
    if base is not None:
        base = _f32c(base, "base")
        if base.shape != src.shape:
            raise RuntimeError("scatter_softmax: base must be [E]")
    out = torch.empty_like(src)
    lib = L.load()
    ws = _workspace(lib.rag_scatter_softmax_workspace(dim_size), src.device)
    if src.numel():
        with torch.cuda.device(src.device):
            L.check(lib.rag_scatter_softmax_f32(_p(src), _p(index), src.numel(), dim_size, lo, span, _p(range_dev), _p(base),
                                                mix_a, mix_b, _p(out), _p(ws), ws.numel(), _stream()), "scatter_softmax")
    return out


@scatter_softmax.register_fake
def _(src, index, dim_size, lo=0.0, span=1.0, base=None, mix_a=0.0, mix_b=1.0, range_dev=None):
    return torch.empty_like(src)


# ----------------------------------------------------------------------------- downstream prompt (a9)
@torch.library.custom_op("ragraph::prompt_act", mutates_args=())
def prompt_act(x: Tensor, w: Tensor, act: int = 0) -> Tensor:
    """act(w * x): w [d] or [1,d]; act 0 identity, 1 ELU."""
    _need_cuda(x, w)
    x, w = _f32c(x, "prompt_act"), _f32c(w, "prompt_act").reshape(-1)
    if x.dim() != 2 or w.numel() != x.shape[1]:
        raise RuntimeError(f"prompt_act: x [n,d] and w [d] expected, got {tuple(x.shape)} and {tuple(w.shape)}")
    out = torch.empty_like(x)
    if x.numel():
        with torch.cuda.device(x.device):
            L.check(L.load().rag_prompt_act_f32(_p(x), x.shape[0], x.shape[1], _p(w), act, _p(out), _stream()),
                    "prompt_act")
    return out


@prompt_act.register_fake
def _(x, w, act=0):
    return torch.empty_like(x)


@torch.library.custom_op("ragraph::prototype_scores", mutates_args=())
def prototype_scores(x: Tensor, proto: Tensor, mode: int = 0, w: Optional[Tensor] = None, act: int = 0,
                     eps: float = 1e-8) -> Tensor:
    """[n,C] cosine of every (optionally prompted) row of x against the class prototypes; mode 0 raw, 1 softmax,
    2 log_softmax over the classes."""
    _need_cuda(x, proto, w)
    x, proto = _f32c(x, "prototype_scores"), _f32c(proto, "prototype_scores")
    if x.dim() != 2 or proto.dim() != 2 or proto.shape[1] != x.shape[1]:
        raise RuntimeError(f"prototype_scores: x [n,d] and proto [C,d] expected, got {tuple(x.shape)}, {tuple(proto.shape)}")
    if w is not None:
        w = _f32c(w, "w").reshape(-1)
        if w.numel() != x.shape[1]:
            raise RuntimeError("prototype_scores: w must have d entries")
    n, d, C_ = x.shape[0], x.shape[1], proto.shape[0]
    out = torch.empty((n, C_), dtype=torch.float32, device=x.device)
    if n:
        with torch.cuda.device(x.device):
            L.check(L.load().rag_prototype_scores_f32(_p(x), n, d, _p(w), act, _p(proto), C_, eps, mode, _p(out),
                                                      _stream()), "prototype_scores")
    return out


@prototype_scores.register_fake
def _(x, proto, mode=0, w=None, act=0, eps=1e-8):
    return x.new_empty((x.shape[0], proto.shape[0]))


def negative_sample(users: Tensor, hist_rowptr: Tensor, hist_items: Tensor, num_items: int, n_neg: int = 1,
                    seed: int = 0) -> Tensor:
    """int64 [M, n_neg]: per (user, slot) one item uniform over [0, num_items) outside the user's history row (CSR of SORTED
    int64 item ids) -- get_train_batch's rejection sampler (RAGraph_edge/utils/dataloader.py:140-152) on the device."""
    _need_cuda(users, hist_rowptr, hist_items)
    if users.dtype != torch.int64 or hist_rowptr.dtype != torch.int64 or hist_items.dtype != torch.int64:
        raise RuntimeError("negative_sample: users, hist_rowptr and hist_items must be int64")
    users, hist_rowptr, hist_items = users.contiguous(), hist_rowptr.contiguous(), hist_items.contiguous()
    M = users.numel()
    out = torch.empty((M, n_neg), dtype=torch.int64, device=users.device)
    with torch.cuda.device(users.device):
        L.check(L.load().rag_negative_sample(_p(users), M, n_neg, _p(hist_rowptr), _p(hist_items), hist_rowptr.numel() - 1,
                                             num_items, seed & 0xFFFFFFFFFFFFFFFF, _p(out), _stream()), "negative_sample")
    return out


def gather_oob_count() -> int:
    return int(L.load().rag_gather_oob_count())
