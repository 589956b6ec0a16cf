"""Mirror of the edge (link-prediction) variant's pieces of the hot path.

scatter_sum / scatter_add : RAGraph_edge/modules/utils.py:17-38
rating_topk               : RAGraph_edge/utils/metrics.py:48-53,96-118 (evaluation ranking)
_agg                      : RAGraph_edge/modules/RAGraph.py:232-240
time encoding             : RAGraph_edge/modules/RAGraph.py:250-263,266-267 (scatter_softmax over destination nodes)
retrieve loop + blend     : RAGraph_edge/modules/RAGraph.py:279-328
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import _lib as L
from . import ops
from .csr import CSRGraph


def scatter_sum(src: Tensor, index: Tensor, dim: int = -1, out: Optional[Tensor] = None,
                dim_size: Optional[int] = None) -> Tensor:
    """out[index[e]] += src[e] along dim 0 (the only form the hot path uses; index is 1-D as in _agg).
    Runs as a CSR SpMM over the identity-column matrix: no atomics on the output."""
    if dim not in (0, -src.dim()) or index.dim() != 1 or src.dim() != 2:
        raise NotImplementedError("scatter_sum: only dim=0 with a 1-D index over a 2-D src is on the hot path")
    E = src.shape[0]
    if dim_size is None:
        dim_size = out.shape[0] if out is not None else (int(index.max()) + 1 if E else 0)
    edges = torch.stack([torch.arange(E, device=src.device), index], dim=1)
    g = CSRGraph.from_coo(edges, None, dim_size, E)
    if out is None:
        return g.spmm(src.contiguous())
    res = g.spmm(src.contiguous(), L.EPI_ACCUM, accum_in=out)
    out.copy_(res)
    return out


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    return scatter_sum(src, index, dim, out, dim_size)


def scatter_softmax(src: Tensor, index: Tensor, dim: int = 0, dim_size: Optional[int] = None) -> Tensor:
    """torch_scatter.scatter_softmax for the form the hot path uses (1-D src, 1-D index): softmax over the entries
    that share an index, result in the caller's order."""
    if src.dim() != 1 or index.dim() != 1 or dim not in (0, -1):
        raise NotImplementedError("scatter_softmax: only 1-D src / index (edge scalars grouped by destination node)")
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    return ops.scatter_softmax(src.float(), index, int(dim_size))


def relative_edge_time_encoding(edges: Tensor, edge_times: Tensor, num_nodes: int, max_step=None,
                                edge_norm: Optional[Tensor] = None) -> Tensor:
    """`_relative_edge_time_encoding` (modules/RAGraph.py:250-263): edge times min-max rescaled to [0,1], then a
    softmax per destination node (edges[:,1]).  The rescale runs inside the kernel; with ``edge_norm`` given the
    result is already the mixed edge weight of :267, ``edge_norm * 1/2 + time_norm * 1/2``."""
    t = edge_times.float()
    hi = t.max() if max_step is None else torch.as_tensor(max_step, dtype=torch.float32, device=t.device)
    rng = torch.stack([t.min(), hi.reshape(())])          # stays on the device: no host sync for the min / max
    if edge_norm is None:
        return ops.scatter_softmax(t, edges[:, 1].contiguous(), num_nodes, range_dev=rng)
    return ops.scatter_softmax(t, edges[:, 1].contiguous(), num_nodes, 0.0, 1.0, edge_norm, 0.5, 0.5, rng)


class EdgeAggregator:
    """`_agg(all_emb, edges, edge_norm)` with the CSR built once per (edges, edge_norm) pair.
    Y[dst] = sum_e w_e * X[src_e]  (src = edges[:,0], dst = edges[:,1])."""

    def __init__(self, num_nodes: int, deterministic: bool = False):
        self.num_nodes = num_nodes
        self.deterministic = deterministic
        self._key = None
        self._refs = None
        self._csr: Optional[CSRGraph] = None

    def csr(self, edges: Tensor, edge_norm: Tensor) -> CSRGraph:
        # The cache entry keeps STRONG references to the tensors it was built from: a key made of addresses alone would
        # match a different tensor that the caching allocator placed at a recycled address (a per-forward time-encoded
        # edge_norm, equal-length edge-dropout samples) and silently reuse a stale CSR.
        key = (edges.data_ptr(), edges._version, edges.shape[0], edge_norm.data_ptr(), edge_norm._version)
        if key != self._key or self._refs is None or self._refs[0] is not edges or self._refs[1] is not edge_norm:
            self._csr = CSRGraph.from_coo(edges, edge_norm, self.num_nodes, self.num_nodes, self.deterministic)
            self._key = key
            self._refs = (edges, edge_norm)
        return self._csr

    def invalidate(self) -> None:
        self._key = self._csr = self._refs = None

    def __call__(self, all_emb: Tensor, edges: Tensor, edge_norm: Tensor, **epi) -> Tensor:
        return self.csr(edges, edge_norm).spmm(all_emb, **epi)


def make_resource_graph(all_emb: Tensor, edges: Tensor, edge_norm: Tensor, radius: int = 3,
                        aggregator: Optional[EdgeAggregator] = None, num_augment_scale: int = 0,
                        num_inverse_sample: int = 0, adj=None):
    """``_make_resource_graph`` (modules/RAGraph.py:185-226): keys = A^radius X0 (the last propagation layer), values =
    sum of the even layers X0 + A^2 X0 + ... (``res_emb[0::2]``); each layer is one SpMM over the CSR built once from
    (edges, edge_norm), no [E, d] temporaries.  With both knobs at 0 (the finetune-phase settings, :45-50) the library is
    every node.  ``num_augment_scale`` adds that many noisy copies (Augmentation.augment_features on keys and values) and
    ``num_inverse_sample`` keeps that many rows per copy, drawn by inverse importance of ``adj`` (the vanilla-phase
    settings, :38-43); both use the reference's draw order, so a seeded call gives the reference's library.
    Returns (resource_keys, resource_values)."""
    agg = aggregator or EdgeAggregator(all_emb.shape[0])
    g = agg.csr(edges, edge_norm)
    layer, values = all_emb, all_emb
    with torch.no_grad():
        for l in range(1, radius + 1):
            layer = g.spmm(layer)
            if l % 2 == 0:
                values = values + layer
    if num_augment_scale == 0 and num_inverse_sample == 0:
        return layer, values
    from .ragraph_utils.Augmentation import Augmentation
    from .sampling import InverseSampling
    # adj[i, j] of the reference is the weight of edge (i, j) = edges row (i, j): the CSR above is grouped by edges[:, 1],
    # i.e. it is adj^T -- the importance scores need adj itself
    sample_prob = InverseSampling.compute_sample_prob(adj if adj is not None else g.transpose())
    keys_out, values_out = [], []
    for i in range(1 + int(num_augment_scale)):
        k_i, v_i = (layer, values) if i == 0 else (Augmentation.augment_features(layer, sample_prob),
                                                   Augmentation.augment_features(values, sample_prob))
        if num_inverse_sample > 0:
            sample_mask = torch.multinomial(sample_prob, num_samples=int(num_inverse_sample), replacement=True)
            k_i, v_i = k_i[sample_mask], v_i[sample_mask]
        keys_out.append(k_i); values_out.append(v_i)
    return torch.cat(keys_out, dim=0), torch.cat(values_out, dim=0)


def _agg(all_emb: Tensor, edges: Tensor, edge_norm: Tensor, num_nodes: int) -> Tensor:
    """Stateless form with the reference argument order (+ num_users+num_items made explicit)."""
    return CSRGraph.from_coo(edges, edge_norm, num_nodes, num_nodes).spmm(all_emb)


def edge_rag_forward(all_emb: Tensor, edges: Tensor, edge_norm: Tensor, resource_keys: Tensor,
                     resource_values: Tensor, num_layers: int = 3, retrieve_num: int = 10, batch_size: int = 4096,
                     retrieve_weight: float = 0.3, key_inv_norm: Optional[Tensor] = None,
                     keys_bf16: Optional[Tensor] = None, mode: int = L.SIM_FP32,
                     aggregator: Optional[EdgeAggregator] = None, edge_times: Optional[Tensor] = None,
                     max_time_step=None, add_noise: bool = False, noise_retrieve_num: int = 1) -> Tensor:
    """modules/RAGraph.py:265-328 behind the LoRA / gating of the embedding tables (dense torch modules that produce
    ``all_emb``): time-aware edge weights when ``edge_times`` is given (:266-267), LightGCN layer sum, retrieval blend.
    res = (1-w) * (X0 + A X0 + A^2 X0 + A^3 X0) + w * mean_k values[topk(cos(X0, keys))].
    Per 4096-query batch the retrieve is one fused similarity+top-k launch and the mean + convex blend one
    gather_reduce launch.  ``add_noise`` (use_noise and training, :296): top-(k + noise_retrieve_num) plus
    noise_retrieve_num uniformly random library rows per query (CPU ``torch.randint`` like :317, same RNG stream),
    all averaged together."""
    n = all_emb.shape[0]
    if edge_times is not None:
        edge_norm = relative_edge_time_encoding(edges, edge_times, n, max_time_step, edge_norm)
    agg = aggregator or EdgeAggregator(n)
    g = agg.csr(edges, edge_norm)
    train = torch.is_grad_enabled() and all_emb.requires_grad
    layer, total = all_emb, all_emb
    for _ in range(num_layers):
        layer = g.spmm(layer)               # differentiable: backward is the same kernel on the cached transpose
        total = total + layer
    if key_inv_norm is None:
        key_inv_norm = ops.row_inv_norm(resource_keys)
    queries = all_emb.detach()              # indices are not differentiable; the library carries no grad (:186-226)
    out = torch.empty_like(queries)
    cosine_topk, gather_reduce = ops.direct(ops.cosine_topk), ops.direct(ops.gather_reduce)   # ~60 batches per forward
    k_eff = retrieve_num + noise_retrieve_num if add_noise else retrieve_num
    for start in range(0, n, batch_size):
        end = min(start + batch_size, n)
        _, idx = cosine_topk(queries[start:end], resource_keys, k_eff, key_inv_norm, keys_bf16, mode)
        if add_noise:
            noise_indices = torch.randint(0, resource_values.shape[0], (end - start, noise_retrieve_num))
            idx = torch.cat([idx, noise_indices.to(idx.device)], dim=1)
        if train:
            out[start:end] = gather_reduce(resource_values, idx, L.REDUCE_MEAN)
        else:
            out[start:end] = gather_reduce(resource_values, idx, L.REDUCE_MEAN, total[start:end], retrieve_weight)
    if train:                               # gradient reaches all_emb through the layer sum only (:327-328)
        return (1 - retrieve_weight) * total + retrieve_weight * out
    return out


_rating_shadow_cache: dict = {}


def _rating_shadow(item_emb: Tensor):
    """(fp16 shadow of item_emb * key_scale, err_max, key_scale) for the tensor-core ranking, cached per tensor object +
    version: an evaluation ranks every user batch against the same item table (utils/metrics.py:96-118).  key_scale =
    1 / (largest item norm) brings every row norm to <= 1, which is what the filter's error bound assumes."""
    key = id(item_emb)
    hit = _rating_shadow_cache.get(key)
    if hit is not None and hit[0]() is item_emb and hit[1] == item_emb._version:
        return hit[2]
    norm_max = float(item_emb.detach().float().norm(dim=1).max())          # one host sync per item table
    scale = 1.0 / norm_max if norm_max > 0.0 else 1.0
    err = torch.zeros(1, dtype=torch.float32, device=item_emb.device)
    shadow, _ = ops.rows_to_shadow16((item_emb.detach().float() * scale).contiguous(), L.FMT_F16, False, err_max=err)
    out = (shadow, err, scale)
    if len(_rating_shadow_cache) > 8:
        _rating_shadow_cache.clear()
    try:
        import weakref
        _rating_shadow_cache[key] = (weakref.ref(item_emb), item_emb._version, out)
    except TypeError:
        pass
    return out


RATING_TC_MIN_ITEMS, RATING_TC_MIN_USERS = 4096, 16


def rating_topk(user_emb: Tensor, item_emb: Tensor, k: int, hist_rowptr: Tensor, hist_items: Tensor,
                tensor_cores: Optional[bool] = None) -> Tensor:
    """Evaluation ranking of RAGraph_edge/utils/metrics.py:96-118 in one op: top-k items by user . item, a user's
    history items excluded (Metric._mask_history_pos sets them to -inf, :48-53).  hist_rowptr int64[B+1] / hist_items
    int64[nnz] hold the batch users' history lists back to back.  Returns item indices int64 [B, k] like
    ``torch.topk(batch_pred, k)[1]``; the [B, n_items] rating matrix and its .cpu() copy never exist.
    Large tables (>= 4 096 items, >= 16 users, d <= 128, k <= 128) run on the tensor cores: fp16 filter over the scaled item
    table + fp32 refine that drops the history items + certificate -- the same ranking as the fp32 kernel
    (``tensor_cores`` = True / False forces one path)."""
    B, d = user_emb.shape
    n_items = item_emb.shape[0]
    tc = tensor_cores
    if tc is None:
        tc = (n_items >= RATING_TC_MIN_ITEMS and B >= RATING_TC_MIN_USERS and user_emb.is_cuda
              and bool(L.load().rag_sim_mode_supported(L.SIM_F16_REFINE, d, k)))
    if not tc:
        return ops.topk_masked(user_emb, item_emb, k, hist_rowptr, hist_items, L.SIM_DOT)[1]
    shadow, err, scale = _rating_shadow(item_emb)
    return ops.topk_masked(user_emb, item_emb, k, hist_rowptr, hist_items, L.SIM_DOT, None, 0, L.SIM_F16_REFINE, shadow, err,
                           scale)[1]


def history_csr(train_user_dict, num_users: int, device) -> tuple:
    """{user: [items]} (the reference's ``train_user_dict``, utils/dataloader.py:62-75) -> (hist_rowptr int64 [U+1],
    hist_items int64 [nnz] sorted within each user) on ``device``: the form ``negative_sampling`` and ``rating_topk`` take."""
    counts = torch.zeros(num_users, dtype=torch.int64)
    rows = []
    for u in range(num_users):
        items = sorted(set(int(i) for i in train_user_dict.get(u, ())))
        counts[u] = len(items)
        rows.extend(items)
    rowptr = torch.zeros(num_users + 1, dtype=torch.int64)
    torch.cumsum(counts, 0, out=rowptr[1:])
    return rowptr.to(device), torch.tensor(rows, dtype=torch.int64).to(device)


def negative_sampling(users: Tensor, hist_rowptr: Tensor, hist_items: Tensor, num_items: int, n: int = 1,
                      seed: Optional[int] = None) -> Tensor:
    """``negative_sampling(user_item, train_user_set, n)`` of get_train_batch (RAGraph_edge/utils/dataloader.py:140-152) on the
    device: for every user of the batch, n items uniform over the items outside the user's training history -- one kernel
    instead of a Python rejection loop per pair.  Returns int64 [M * n] in the reference's order (user-major).  ``seed``
    defaults to a draw from torch's CPU generator, so ``torch.manual_seed`` makes a run reproducible."""
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    return ops.negative_sample(users, hist_rowptr, hist_items, num_items, n, seed).reshape(-1)
