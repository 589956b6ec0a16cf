"""CPU oracle for the RAGraph retrieve -> gather -> propagate hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline.  Nothing under
``ragraph_b200/`` imports this module; the product path fails loudly when the CUDA
library is missing.

Parity pinning: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md section 4 / 8c), so the oracle is pinned against outputs of the reference
ITSELF: ``oracle/make_golden.py`` imports the unmodified reference functions from
``/root/reference`` (in the build container, where it exists), runs them on seeded
inputs and commits the input/output vectors under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every function below against those vectors.

The arithmetic of the reference lives in PyTorch (``F.normalize``, ``torch.matmul``,
``torch.topk``, advanced indexing, ``scatter_add_``); the restatement therefore uses the
same torch CPU calls in the same order (fp32), plus numpy fp64 arbiters used to decide
near-ties.  Every function cites the reference lines it follows (paths relative to
``/root/reference``).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-12  # F.normalize default eps


# ----------------------------------------------------------------------------------------
# a1  SimilarityFunctions.calculate_cosine_similarity
# ----------------------------------------------------------------------------------------
def cosine_similarity(search_keys: torch.Tensor, resource_keys: torch.Tensor) -> torch.Tensor:
    """RAGraph_node/ragraph_utils/SimilarityFunctions.py:6-16 (identical in all 5 variants).

    x / max(||x||_2, 1e-12) on BOTH operands (keys are re-normalised on every call),
    then a dense fp32 matmul.  A 1-D query (graph variant, RAGraph_graph/RAGraph.py:50)
    gives a 1-D result.
    """
    q = F.normalize(search_keys, p=2, dim=-1)
    k = F.normalize(resource_keys, p=2, dim=-1)
    return torch.matmul(q, k.t())


def cosine_similarity_f64(search_keys, resource_keys) -> np.ndarray:
    """fp64 arbiter for a1 (same formula, numpy float64)."""
    q = np.asarray(search_keys, dtype=np.float64)
    k = np.asarray(resource_keys, dtype=np.float64)
    qn = q / np.maximum(np.linalg.norm(q, axis=-1, keepdims=True), EPS)
    kn = k / np.maximum(np.linalg.norm(k, axis=-1, keepdims=True), EPS)
    return qn @ kn.T


def dot_similarity(search_keys: torch.Tensor, resource_keys: torch.Tensor) -> torch.Tensor:
    """Un-normalised score used by the edge evaluation top-k,
    RAGraph_edge/utils/metrics.py:102-117 (rating = U . I^T)."""
    return torch.matmul(search_keys, resource_keys.t())


# ----------------------------------------------------------------------------------------
# a2 / a3  ToyGraphBase.retrieve  (top-k + gathers)
# ----------------------------------------------------------------------------------------
def topk(scores: torch.Tensor, k: int):
    """RAGraph_node/ragraph_utils/ToyGraphBase.py:67 -- torch.topk(largest, sorted)."""
    return torch.topk(scores, k, largest=True, sorted=True)


def retrieve(search_keys, resource_keys, resource_values, resource_labels, retrieve_num,
             noise_indices=None):
    """RAGraph_node/ragraph_utils/ToyGraphBase.py:47-81.

    Returns (topk_scores, topk_indices, rag_embeddings, rag_labels).  ``noise_indices``
    ([Q, noise_retrieve_num] int64) stands in for the CPU ``torch.randint`` of :74 so that
    the noisy branch is reproducible; the caller passes ``retrieve_num`` already doubled
    when add_noise is set (:66).
    """
    scores = cosine_similarity(search_keys, resource_keys)
    topk_scores, topk_indices = topk(scores, retrieve_num)
    rag_embeddings = resource_values[topk_indices]
    rag_labels = resource_labels[topk_indices]
    if noise_indices is not None:
        rag_embeddings = torch.cat([rag_embeddings, resource_values[noise_indices]], dim=1)
        rag_labels = torch.cat([rag_labels, resource_labels[noise_indices]], dim=1)
    return topk_scores, topk_indices, rag_embeddings, rag_labels


def retrieve_graph(search_key_1d, resource_keys, resource_values, resource_labels, retrieve_num):
    """RAGraph_graph/ragraph_utils/ToyGraphBase.py:56-87: the query is ONE vector [d]
    (mean of node embeddings), similarity is [N] and is unsqueezed to [1, N] (:72)."""
    scores = cosine_similarity(search_key_1d, resource_keys).unsqueeze(0)
    topk_scores, topk_indices = topk(scores, retrieve_num)
    return topk_scores, topk_indices, resource_values[topk_indices], resource_labels[topk_indices]


def retrieve_two_metric(search_keys, search_positions, resource_keys, resource_positions,
                        resource_values, resource_labels, retrieve_num,
                        structure_weight=0.001, semantic_weight=0.999):
    """RAGraph_node_fewshot/ragraph_utils/ToyGraphBase.py:47-79: weighted sum of two cosine
    matrices (einsum 'ij,jkl->ikl' over a [1,2] weight row, :56-61) before the top-k."""
    structure = cosine_similarity(search_positions, resource_positions)
    semantic = cosine_similarity(search_keys, resource_keys)
    w = torch.tensor([[structure_weight, semantic_weight]], dtype=semantic.dtype)
    mats = torch.stack([structure, semantic], dim=0)
    scores = torch.einsum('ij,jkl->ikl', w, mats).squeeze(0)
    topk_scores, topk_indices = topk(scores, retrieve_num)
    return topk_scores, topk_indices, resource_values[topk_indices], resource_labels[topk_indices]


def gather_rows(table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """ToyGraphBase.py:70-71 / RAGraph_edge/modules/RAGraph.py:314 -- ``table[idx]``."""
    return table[idx]


# ----------------------------------------------------------------------------------------
# a4  Propagation.aggregate_k_hop_features
# ----------------------------------------------------------------------------------------
def aggregate_k_hop_features(adj: torch.Tensor, x: torch.Tensor, k: int) -> torch.Tensor:
    """RAGraph_node/ragraph_utils/Propagation.py:7-27: row-normalise the dense adjacency by
    its row sums (:15-16), then k x relu(A_norm @ x) (:19-25).  k = 0 returns x."""
    out = x
    degree = adj.sum(dim=1, keepdim=True)
    adj_normalized = adj / degree
    for _ in range(k):
        out = torch.matmul(adj_normalized, out)
        out = F.relu(out)
    return out


def aggregate_k_hop_features_f64(adj, x, k) -> np.ndarray:
    a = np.asarray(adj, dtype=np.float64)
    out = np.asarray(x, dtype=np.float64)
    a = a / a.sum(axis=1, keepdims=True)
    for _ in range(k):
        out = np.maximum(a @ out, 0.0)
    return out


# ----------------------------------------------------------------------------------------
# a5  GCN.forward
# ----------------------------------------------------------------------------------------
def gcn_layer(seq, adj, weight, bias, prelu_alpha) -> torch.Tensor:
    """RAGraph_node/layers/gcn.py:26-40 (dense branch): PReLU(adj @ (seq @ W^T) + b).
    ``weight`` is nn.Linear(in, out, bias=False).weight ([out, in]); ``prelu_alpha`` the
    single nn.PReLU() parameter (:9).  adj may be [1, n, n] (squeezed at :36)."""
    seq_fts = F.linear(seq, weight)
    out = torch.mm(adj.squeeze(dim=0), seq_fts)
    if bias is not None:
        out = out + bias
    alpha = torch.as_tensor(prelu_alpha, dtype=out.dtype).reshape(-1)
    return F.prelu(out, alpha)


# ----------------------------------------------------------------------------------------
# a6  edge _agg + scatter_sum
# ----------------------------------------------------------------------------------------
def scatter_sum(src: torch.Tensor, index: torch.Tensor, dim: int = 0, dim_size=None) -> torch.Tensor:
    """RAGraph_edge/modules/utils.py:17-32 (dim=0 form used by _agg): zeros + scatter_add_."""
    assert dim == 0
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype)
    idx = index.reshape(-1, *([1] * (src.dim() - 1))).expand_as(src)
    return out.scatter_add_(0, idx, src)


def edge_agg(all_emb: torch.Tensor, edges: torch.Tensor, edge_norm: torch.Tensor, num_nodes: int):
    """RAGraph_edge/modules/RAGraph.py:232-240: Y[dst] += w_e * X[src];
    src = edges[:, 0], dst = edges[:, 1]."""
    src_emb = all_emb[edges[:, 0]]
    src_emb = src_emb * edge_norm.unsqueeze(1)
    return scatter_sum(src_emb, edges[:, 1], dim=0, dim_size=num_nodes)


def edge_agg_f64(all_emb, edges, edge_norm, num_nodes) -> np.ndarray:
    x = np.asarray(all_emb, dtype=np.float64)
    e = np.asarray(edges)
    w = np.asarray(edge_norm, dtype=np.float64)
    out = np.zeros((num_nodes, x.shape[1]), dtype=np.float64)
    np.add.at(out, e[:, 1], x[e[:, 0]] * w[:, None])
    return out


# ----------------------------------------------------------------------------------------
# a7  fusion arithmetic of RAGraph.forward
# ----------------------------------------------------------------------------------------
def task_decoder(x, w1, b1, w2, b2):
    """RAGraph_node/ragraph_utils/TaskDecoder.py:3-17: Linear -> LeakyReLU(0.01) -> Linear."""
    return F.linear(F.leaky_relu(F.linear(x, w1, b1), 0.01), w2, b2)


def fuse_node(pretrain_emb, adj, rag_embeddings, rag_labels, decoder_params,
              retrieve_weight=0.5, label_weight=0.5, hop=3, finetune=True):
    """RAGraph_node/RAGraph.py:47-63."""
    rag_label = torch.mean(rag_labels, dim=1)
    if not finetune:
        return rag_label
    rag_embedding = torch.sum(rag_embeddings, dim=1)
    query_embeddings = aggregate_k_hop_features(adj, pretrain_emb, hop)
    hidden = query_embeddings * (1 - retrieve_weight) + rag_embedding * retrieve_weight
    decode = torch.softmax(task_decoder(hidden, *decoder_params), dim=1)
    return decode * (1 - label_weight) + rag_label * label_weight


def fuse_graph(pretrain_emb, adj, rag_embeddings, rag_labels, decoder_params,
               retrieve_weight=0.3, label_weight=0.3, hop=1, finetune=True):
    """RAGraph_graph/RAGraph.py:58-75 (one query = mean over the graph's nodes)."""
    rag_label = torch.mean(rag_labels.to(torch.float32), dim=1)
    if not finetune:
        return rag_label
    rag_embedding = torch.sum(rag_embeddings, dim=1)
    query_embedding = torch.mean(aggregate_k_hop_features(adj, pretrain_emb, hop), dim=0)
    hidden = query_embedding * (1 - retrieve_weight) + rag_embedding * retrieve_weight
    decode = torch.softmax(task_decoder(hidden, *decoder_params), dim=1)
    return decode * (1 - label_weight) + rag_label * label_weight


def edge_forward(all_emb, edges, edge_norm, resource_keys, resource_values, num_layers=3,
                 retrieve_num=10, batch_size=4096, retrieve_weight=0.3, noise_indices=None, noise_retrieve_num=1):
    """RAGraph_edge/modules/RAGraph.py:279-328 (no LoRA/gating): LightGCN layers via
    _agg, then per 4096-query batch cosine -> top-k -> values[idx].mean(1), then the blend
    res = (1-w) * sum(layers) + w * rag_emb.  ``noise_indices`` ([n, noise_retrieve_num], the
    per-batch torch.randint draws of :317 stacked) switches the noisy branch on: top-(k + noise)
    plus the random rows, all averaged (:310,316-322)."""
    n = all_emb.shape[0]
    res_emb = [all_emb]
    for _ in range(num_layers):
        res_emb.append(edge_agg(res_emb[-1], edges, edge_norm, n))
    query_emb = res_emb[0]
    rag_emb = torch.empty((n, resource_values.shape[1]), dtype=query_emb.dtype)
    for start in range(0, n, batch_size):
        end = min(start + batch_size, n)
        scores = cosine_similarity(query_emb[start:end], resource_keys)
        _, idx = topk(scores, retrieve_num + noise_retrieve_num if noise_indices is not None else retrieve_num)
        batch_rag_emb = resource_values[idx]
        if noise_indices is not None:
            batch_rag_emb = torch.cat([batch_rag_emb, resource_values[noise_indices[start:end]]], dim=1)
        rag_emb[start:end] = batch_rag_emb.mean(dim=1)
    res = sum(res_emb)
    return (1 - retrieve_weight) * res + retrieve_weight * rag_emb


# ----------------------------------------------------------------------------------------
# K10  edge time encoding (scatter_softmax is third-party torch_scatter 2.1.2, absent here:
#      restated from its documented semantics -- softmax over the entries sharing an index)
# ----------------------------------------------------------------------------------------
def scatter_softmax(src: torch.Tensor, index: torch.Tensor, dim_size=None) -> torch.Tensor:
    """torch_scatter.scatter_softmax(src, index, dim_size=...) as called at
    RAGraph_edge/modules/RAGraph.py:261 (1-D src, 1-D index): exp(src - groupmax) / groupsum."""
    n = int(dim_size) if dim_size is not None else int(index.max()) + 1
    mx = torch.full((n,), -float("inf"), dtype=src.dtype).scatter_reduce(0, index, src, reduce="amax", include_self=True)
    e = torch.exp(src - mx[index])
    den = torch.zeros(n, dtype=src.dtype).scatter_add_(0, index, e)
    return e / den[index]


def relative_edge_time_encoding(edges: torch.Tensor, edge_times: torch.Tensor, num_nodes: int, max_step=None):
    """RAGraph_edge/modules/RAGraph.py:250-263: min-max rescale to [0,1], softmax per destination node."""
    edge_times = edge_times.float()
    if max_step is None:
        max_step = edge_times.max()
    edge_times = (edge_times - edge_times.min()) / (max_step - edge_times.min())
    return scatter_softmax(edge_times, edges[:, 1], dim_size=num_nodes)


def edge_time_mix(edge_norm: torch.Tensor, time_norm: torch.Tensor) -> torch.Tensor:
    """RAGraph_edge/modules/RAGraph.py:267."""
    return edge_norm * 1 / 2 + time_norm * 1 / 2


# ----------------------------------------------------------------------------------------
# a9  downstream prompt (downprompt.py)
# ----------------------------------------------------------------------------------------
def downstream_prompt(graph_embedding: torch.Tensor, weight: torch.Tensor, elu: bool) -> torch.Tensor:
    """RAGraph_node/downprompt.py:118-130 (weight * emb, then ELU) / RAGraph_graph/downprompt.py:197-209 (no ELU)."""
    out = weight * graph_embedding
    return F.elu(out) if elu else out


def average_emb(labels: torch.Tensor, rawret: torch.Tensor, nb_class: int, slots: int) -> torch.Tensor:
    """RAGraph_node/downprompt.py:59-79 (nb_class 3, slots = n // 2) / RAGraph_graph/downprompt.py:59-94 (slots = n):
    rows of class c are packed into scratch[c, 0:cnt_c], then torch.mean over the slot axis.  The reference allocates
    the scratch uninitialised (torch.FloatTensor(...)); its result is defined only if the unwritten slots are zero,
    which is what is restated (and what oracle/make_golden.py pins by zero-filling that allocation)."""
    retlabel = torch.zeros(nb_class, slots, rawret.shape[1])
    cnt = [0] * nb_class
    for x in range(rawret.shape[0]):
        c = int(labels[x].item())
        if 0 <= c < nb_class:
            retlabel[c][cnt[c]] = rawret[x]
            cnt[c] += 1
    return torch.mean(retlabel, dim=1)


def prototype_scores(rawret: torch.Tensor, ave: torch.Tensor, mode: str = "raw") -> torch.Tensor:
    """RAGraph_node/downprompt.py:36-46 (softmax) / RAGraph_graph/downprompt.py:41-56 (log_softmax): one
    torch.cosine_similarity(rawret[x], ave[c], dim=0) per (row, class)."""
    ret = torch.empty(rawret.shape[0], ave.shape[0])
    for x in range(rawret.shape[0]):
        for c in range(ave.shape[0]):
            ret[x][c] = torch.cosine_similarity(rawret[x], ave[c], dim=0)
    if mode == "softmax":
        return F.softmax(ret, dim=1)
    if mode == "log_softmax":
        return F.log_softmax(ret, dim=1)
    return ret


def split_and_batchify_graph_feats(batched_graph_feats: torch.Tensor, graph_sizes: torch.Tensor) -> torch.Tensor:
    """RAGraph_graph/downprompt.py:98-112: per-graph sum over consecutive row blocks."""
    cnt = 0
    result = torch.zeros(graph_sizes.shape[0], batched_graph_feats.shape[1])
    for i in range(graph_sizes.shape[0]):
        n_i = int(graph_sizes[i].item())
        result[i] = torch.sum(batched_graph_feats[cnt:cnt + n_i, :], dim=0)
        cnt += n_i
    return result


# ----------------------------------------------------------------------------------------
# a7  few-shot fusion (labels -> class id -> mean few-shot logits; decode = second GCN layer)
# ----------------------------------------------------------------------------------------
def fuse_fewshot(pretrain_emb, adj, rag_embeddings, rag_labels, mean_fewshot_logits, decode_params,
                 retrieve_weight, label_weight, hop, finetune=True, graph_level=False):
    """RAGraph_node_fewshot/RAGraph.py:47-83 (graph_level=False) and RAGraph_graph_fewshot/RAGraph.py:46-91
    (graph_level=True: every node of the single query graph retrieves, the blended logits are averaged over nodes).
    decode_params = (weight, bias, prelu_alpha) of the second GCN layer (models/gcnlayers.py `decode`)."""
    rag_logits = mean_fewshot_logits[torch.argmax(rag_labels, dim=-1)]
    rag_logits = torch.mean(rag_logits, dim=1)
    if not finetune:
        return rag_logits
    rag_embedding = torch.sum(rag_embeddings, dim=1)
    query_embeddings = aggregate_k_hop_features(adj, pretrain_emb, hop)
    hidden = query_embeddings * (1 - retrieve_weight) + rag_embedding * retrieve_weight
    decode_logits = gcn_layer(hidden, adj, *decode_params)
    label_logits = decode_logits * (1 - label_weight) + rag_logits * label_weight
    if graph_level:
        label_logits = label_logits.mean(dim=0).unsqueeze(0)
    return label_logits


# ----------------------------------------------------------------------------------------
# a8 / 8f-2  library construction (deterministic core: no augmentation, no inverse sampling)
# ----------------------------------------------------------------------------------------
def build_library_rows_node(embeddings, adj, node_labels, toy_graph_hop):
    """RAGraph_node/ragraph_utils/ToyGraphBase.py:104-119 for one graph (the `else` branch of :95-107):
    keys = F.normalize(emb), values = aggregate_k_hop_features(adj, keys, hop - 1), labels = node labels."""
    keys = F.normalize(embeddings, p=2, dim=-1)
    values = aggregate_k_hop_features(adj, keys, toy_graph_hop)
    return keys, values, node_labels


def build_library_rows_graph(embeddings, adj, graph_label, num_class, toy_graph_hop):
    """RAGraph_graph/ragraph_utils/ToyGraphBase.py:112-126: one row per graph -- node keys / values averaged,
    label = one_hot(graph_label) (int64)."""
    keys = F.normalize(embeddings, p=2, dim=-1)
    values = aggregate_k_hop_features(adj, keys, toy_graph_hop)
    return (torch.mean(keys, dim=0).unsqueeze(0), torch.mean(values, dim=0).unsqueeze(0),
            F.one_hot(graph_label, num_classes=num_class))


def edge_resource_graph(all_emb, edges, edge_norm, radius):
    """RAGraph_edge/modules/RAGraph.py:185-196,212-226 with num_augment_scale = num_inverse_sample = 0:
    keys = last propagation layer, values = sum of the even layers (res_emb[0::2])."""
    n = all_emb.shape[0]
    res_emb = [all_emb]
    for _ in range(radius):
        res_emb.append(edge_agg(res_emb[-1], edges, edge_norm, n))
    return res_emb[-1], sum(res_emb[0::2])


# ----------------------------------------------------------------------------------------
# 8f-2  inverse-importance sampling probabilities (library build; PageRank as an SpMV)
# ----------------------------------------------------------------------------------------
def pagerank(adj: torch.Tensor, d=0.85, eps=1e-6) -> torch.Tensor:
    """RAGraph_node/ragraph_utils/InverseSampling.py:22-47 (dense): row-normalised transition matrix, rows without
    out-links replaced by the uniform row, power iteration until the L1 change is < eps; returns the PREVIOUS iterate."""
    N = adj.shape[0]
    out_degree = torch.sum(adj, dim=1)
    zero_out_degree = out_degree == 0
    out_degree[zero_out_degree] = 1
    adj_normalized = adj / out_degree[:, None]
    adj_normalized[zero_out_degree] = 1.0 / N
    p = torch.ones(N, dtype=torch.float32) / N
    adj_normalized_t = adj_normalized.t()
    while True:
        new_p = (1 - d) / N + d * torch.mv(adj_normalized_t, p)
        if torch.norm(new_p - p, p=1) < eps:
            break
        p = new_p
    return p


def degree_centrality(adj: torch.Tensor) -> torch.Tensor:
    """InverseSampling.py:49-56: column sums / (N - 1)."""
    return torch.sum(adj, dim=0) / (adj.shape[0] - 1)


def sample_prob(adj: torch.Tensor) -> torch.Tensor:
    """InverseSampling.py:6-19: probabilities proportional to 1 / (0.5 PageRank + 0.5 degree centrality + 1e-6)."""
    node_importance = 0.5 * pagerank(adj.clone()) + 0.5 * degree_centrality(adj)
    inverse_node_importance = 1 / (node_importance + 1e-6)
    return inverse_node_importance / torch.sum(inverse_node_importance)


# ----------------------------------------------------------------------------------------
# few-shot helpers around the few-shot forward (RAGraph_node_fewshot/ragraph_utils/utility.py:75-162)
# ----------------------------------------------------------------------------------------
def fewshot_mean(fewshot_logits: torch.Tensor, fewshot_labels: torch.Tensor):
    """utility.py:75-92: mean logits per unique label (sorted), one boolean mask per label."""
    unique_labels = fewshot_labels.unique()
    return torch.stack([fewshot_logits[fewshot_labels == label].mean(dim=0) for label in unique_labels]), unique_labels


def fewshot_predict_logits(mean_fewshot_logits: torch.Tensor, logits: torch.Tensor) -> torch.Tensor:
    """utility.py:129-134: broadcast cosine similarity [n, C]."""
    return F.cosine_similarity(logits.unsqueeze(1), mean_fewshot_logits.unsqueeze(0), dim=-1)


# ----------------------------------------------------------------------------------------
# multi-GPU restatement (new functionality, C1): merge of per-shard candidates
# ----------------------------------------------------------------------------------------
def rating_topk(user_emb: torch.Tensor, item_emb: torch.Tensor, hist_rowptr, hist_items, k: int):
    """Evaluation ranking, RAGraph_edge/utils/metrics.py:110-117 + 210-214: rating = user_emb @ item_emb.T
    (LightGCN.rating, modules/LightGCN.py:115-116), every history item of a user set to -1e8, torch.topk(k)."""
    pred = torch.matmul(user_emb, item_emb.t()).clone()
    rp = [int(x) for x in hist_rowptr]
    for i in range(pred.shape[0]):
        pos_list = [int(x) for x in hist_items[rp[i]:rp[i + 1]]]
        pred[i, pos_list] = -1e8
    return torch.topk(pred, k=k)


def process_tu_arrays(xs, edge_indices, num_node_attributes: int):
    """process_tu_dataset on plain arrays (RAGraph_node/ragraph_utils/utility.py:30-72): dense block-diagonal adjacency
    via scipy coo -> todense per graph, then normalize_adj(adj + I).todense() (utils/process.py:208-215), float32."""
    import scipy.sparse as sp
    features = rawlabels = adjacency = None
    for g, (x, e_ind) in enumerate(zip(xs, edge_indices)):
        x = np.asarray(x); e_ind = np.asarray(e_ind)
        f, l = x[:, :num_node_attributes], x[:, num_node_attributes:]
        coo = sp.coo_matrix((np.ones(e_ind.shape[1]), (e_ind[0, :], e_ind[1, :])), shape=(x.shape[0], x.shape[0]))
        tmpadj = coo.todense()
        if g == 0:
            features, rawlabels, adjacency = f, l, tmpadj
        else:
            features = np.vstack((features, f)); rawlabels = np.vstack((rawlabels, l))
            zero = np.zeros((adjacency.shape[0], x.shape[0]))
            adjacency = np.vstack((np.column_stack((adjacency, zero)), np.column_stack((zero.T, tmpadj))))
    adj = sp.coo_matrix(sp.csr_matrix(adjacency) + sp.eye(adjacency.shape[0]))
    rowsum = np.array(adj.sum(1))
    d_inv_sqrt = np.power(rowsum, -0.5).flatten()
    d_inv_sqrt[np.isinf(d_inv_sqrt)] = 0.
    d = sp.diags(d_inv_sqrt)
    adj = adj.dot(d).transpose().dot(d).tocoo().todense()
    return torch.FloatTensor(np.asarray(features)), torch.FloatTensor(np.asarray(adj)), torch.FloatTensor(np.asarray(rawlabels))


def merge_topk(scores: torch.Tensor, idx: torch.Tensor, k: int):
    """scores/idx: [R, Q, k_r] per-shard candidates (idx already global).  Returns the
    global top-k with the deterministic order the product uses: score desc, index asc."""
    R, Q, kr = scores.shape
    s = scores.permute(1, 0, 2).reshape(Q, R * kr).numpy()
    i = idx.permute(1, 0, 2).reshape(Q, R * kr).numpy()
    order = np.lexsort((i, -s.astype(np.float64)), axis=1)[:, :k]
    return (torch.from_numpy(np.take_along_axis(s, order, axis=1)),
            torch.from_numpy(np.take_along_axis(i, order, axis=1)))


# ----------------------------------------------------------------------------------------
# tolerance helpers (north_star: index sets identical except ties within 1e-6; scores and
# propagated embeddings within 1e-5 relative in fp32; gathers bit-exact)
# ----------------------------------------------------------------------------------------
def topk_sets_match(got_idx, exact_scores_f64: np.ndarray, k: int, tie_tol: float = 1e-6):
    """True iff every returned index has an fp64 score >= (k-th best fp64 score - tie_tol),
    rows hold no duplicates, and every key whose score beats the k-th best by more than
    tie_tol is present.  This is 'identical index sets except for ties within tie_tol'."""
    got = np.asarray(got_idx)
    S = np.asarray(exact_scores_f64)
    if S.ndim == 1:
        S = S[None]
    Q, N = S.shape
    assert got.shape == (Q, k), (got.shape, (Q, k))
    kth = np.partition(S, N - k, axis=1)[:, N - k]
    bad = []
    for r in range(Q):
        row = got[r]
        if len(set(row.tolist())) != k:
            bad.append((r, "duplicate")); continue
        if np.any(S[r, row] < kth[r] - tie_tol):
            bad.append((r, "below kth")); continue
        must = np.nonzero(S[r] > kth[r] + tie_tol)[0]
        if not set(must.tolist()).issubset(set(row.tolist())):
            bad.append((r, "missing"))
    return len(bad) == 0, bad


def rel_err(got, ref) -> float:
    """max |got-ref| / max(|ref|_inf, tiny): relative to the tensor's scale (a per-element
    relative error is meaningless for entries that are exactly or nearly zero after ReLU)."""
    g = np.asarray(got, dtype=np.float64)
    r = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(g - r)) / max(np.max(np.abs(r)), 1e-30)) if r.size else 0.0


def recall_at_k(got_idx, ref_idx) -> float:
    g = np.asarray(got_idx); r = np.asarray(ref_idx)
    hits = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(g, r))
    return hits / float(r.size)
