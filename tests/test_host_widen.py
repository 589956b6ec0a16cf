"""CPU checks of the HOST logic of the widened rows (few-shot fusion forward, downstream prompt heads, edge time
encoding): the CUDA ops are replaced, for these tests only, by stand-ins computed with the oracle, like the gloo tests
inject the compute callables of the sharded retriever.  What is verified is the plumbing around the kernels -- CSR
construction for class / graph indicators, class-id lookup, reduction and blend order, argument order of the ops --
against the reference-generated golden vectors.  The kernels themselves are covered by tests/test_gpu_widen.py."""
import numpy as np
import pytest
import torch

import ragraph_b200 as R
from ragraph_b200 import _lib as L
from ragraph_b200 import downprompt as DP
from ragraph_b200 import ops
from oracle import ragraph_oracle as O

T = torch.from_numpy


def _spmm(rowptr, col, val, x, epilogue=0, bias=None, alpha=None, blend_in=None, blend_w=0.0, accum_in=None):
    n = rowptr.numel() - 1
    rows = torch.repeat_interleave(torch.arange(n), rowptr[1:] - rowptr[:-1])
    v = torch.ones(col.numel()) if val is None else val
    y = torch.zeros(n, x.shape[1]).index_add_(0, rows, x[col.long()] * v[:, None])
    if epilogue & L.EPI_ROWNORM:
        y = y / torch.zeros(n).index_add_(0, rows, v)[:, None]
    if epilogue & L.EPI_BIAS:
        y = y + bias
    if epilogue & L.EPI_RELU:
        y = torch.relu(y)
    if epilogue & L.EPI_PRELU:
        y = torch.where(y >= 0, y, alpha.reshape(-1)[0] * y)
    if epilogue & L.EPI_BLEND:
        y = (1 - blend_w) * y + blend_w * blend_in
    if epilogue & L.EPI_ACCUM:
        y = y + accum_in
    return y


def _csr_from_dense(adj):
    nz = adj != 0
    rowptr = torch.zeros(adj.shape[0] + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(nz.sum(1), 0)
    r, c = nz.nonzero(as_tuple=True)
    return rowptr, c.to(torch.int32), adj[r, c]


def _csr_from_coo(edges, w, n_rows):
    dst, perm = torch.sort(edges[:, 1], stable=True)
    rowptr = torch.zeros(n_rows + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(dst, minlength=n_rows), 0)
    val = torch.ones(edges.shape[0]) if w is None else w[perm]
    return rowptr, edges[:, 0][perm].to(torch.int32), val


def _gather_reduce(table, idx, op=0, blend_in=None, blend_w=0.0):
    r = table[idx].sum(1) if op == L.REDUCE_SUM else table[idx].mean(1)
    return r if blend_in is None else (1 - blend_w) * blend_in + blend_w * r


def _prototype_scores(x, proto, mode=0, w=None, act=0, eps=1e-8):
    if w is not None:
        x = O.downstream_prompt(x, w.reshape(1, -1), bool(act))
    return O.prototype_scores(x, proto, {0: "raw", 1: "softmax", 2: "log_softmax"}[mode])


def _scatter_softmax(src, index, dim_size, lo=0.0, span=1.0, base=None, mix_a=0.0, mix_b=1.0, range_dev=None):
    if range_dev is not None:
        lo, span = range_dev[0], range_dev[1] - range_dev[0]
    sm = O.scatter_softmax((src - lo) / span, index, dim_size)
    return mix_b * sm if base is None else mix_a * base + mix_b * sm


@pytest.fixture
def cpu_ops(monkeypatch):
    """ops.* -> oracle-backed stand-ins (tests of host logic only)."""
    table = {
        "csr_spmm": _spmm, "csr_from_dense": _csr_from_dense, "csr_from_coo": _csr_from_coo,
        "gather_rows": lambda table, idx: table[idx], "gather_reduce": _gather_reduce,
        "cosine_topk": lambda q, keys, k, inv=None, sh=None, mode=0, flags=0, off=0, err=None: torch.topk(O.cosine_similarity(q, keys), k),
        "cosine2_topk": lambda qa, ka, wa, qb, kb, wb, k: torch.topk(
            wa * O.cosine_similarity(qa, ka) + wb * O.cosine_similarity(qb, kb), k),
        "row_inv_norm": lambda x, eps=1e-12: 1.0 / x.norm(dim=1).clamp_min(eps),
        "prompt_act": lambda x, w, act=0: O.downstream_prompt(x, w.reshape(1, -1), bool(act)),
        "prototype_scores": _prototype_scores, "scatter_softmax": _scatter_softmax,
        "rows_normalize": lambda x, eps=1e-12: torch.nn.functional.normalize(x, p=2, dim=-1, eps=eps),
        "cosine_similarity": lambda q, keys, flags=0: O.cosine_similarity(q, keys),
    }
    for name, fn in table.items():
        monkeypatch.setattr(ops, name, fn)
    return ops


def test_downprompt_node_host_logic(golden, cpu_ops):
    g = golden("downprompt_node")
    d = g["seq"].shape[1]
    p = torch.zeros(1, d)
    m = DP.downprompt(p, p, p, d, 3, T(g["feature"]), T(g["labels"]))
    m.downprompt.weight.data.copy_(T(g["weight"]))
    assert O.rel_err(m.ave, g["ave_init"]) < 1e-6
    assert float((m(T(g["seq"]), 0) - T(g["probs_eval"])).abs().max()) < 1e-6
    assert float((m(T(g["seq"]), 1) - T(g["probs_train"])).abs().max()) < 1e-6
    assert O.rel_err(m.ave, g["ave_train"]) < 1e-6


def test_downprompt_graph_host_logic(golden, cpu_ops):
    g = golden("downprompt_graph")
    d = g["seq"].shape[1]
    p = torch.zeros(1, d)
    m = DP.downprompt(p, p, p, d, 6)
    m.downprompt.weight.data.copy_(T(g["weight"]))
    gemb = m(T(g["seq"]), T(g["graph_sizes"]))
    assert O.rel_err(gemb, g["graph_emb"]) < 1e-6
    ave = DP.averageemb(T(g["graph_labels"]), gemb, 6, slots=gemb.shape[0])
    assert O.rel_err(ave, g["ave"]) < 1e-6
    assert float((DP.predict(gemb.shape[0], 6, gemb, ave) - T(g["log_probs"])).abs().max()) < 1e-6
    # labels outside [0, C) own no prototype row, like the reference's chain of ifs
    lab = T(g["graph_labels"]).clone(); lab[0] = 9
    ave2 = DP.averageemb(lab, gemb, 6, slots=gemb.shape[0])
    ref2 = O.average_emb(lab, gemb, 6, gemb.shape[0])
    assert O.rel_err(ave2, ref2) < 1e-6
    with pytest.raises(RuntimeError, match="graph sizes"):
        DP.split_and_batchify_graph_feats(T(g["seq"]), T(g["graph_sizes"]) * 2)


@pytest.mark.parametrize("graph_level", [False, True])
def test_fewshot_forward_host_logic(golden, cpu_ops, graph_level):
    g = golden("fewshot_forward_graph" if graph_level else "fewshot_forward_node")
    d, C = g["keys"].shape[1], g["labels"].shape[1]

    class Backbone:
        def __init__(self):
            self.layer = R.GCN(d, g["dec_weight"].shape[0], "prelu")
            with torch.no_grad():
                self.layer.fc.weight.copy_(T(g["dec_weight"]))
                self.layer.bias.copy_(T(g["dec_bias"]))
                self.layer.act.weight.copy_(T(g["dec_alpha"]))

        def encode(self, features, adj):
            return T(g["emb_q"])

        def decode(self, hidden, adj):
            return self.layer((hidden, adj))

    base = R.ToyGraphBase(None, C, d, int(g["hop"]), device="cpu", variant="node" if graph_level else "node_fewshot")
    base.retrieve_num = int(g["retrieve_num"])
    base.add_entries(T(g["keys"]), T(g["values"]), T(g["labels"]), None if graph_level else T(g["positions"]))
    m = R.RAGraphFewShot(Backbone(), base, d, True, False, int(g["hop"]), float(g["retrieve_weight"]),
                         float(g["label_weight"]), graph_level).eval()
    sp = None if graph_level else T(g["search_positions"])
    adj, logits = T(g["adj"]), T(g["mean_fewshot_logits"])
    with torch.no_grad():
        out = m(None, adj, logits, sp)
        m.finetune = False
        van = m(None, adj, logits, sp)
        m.finetune, m.noise_finetune = True, True
        m.train()
        noisy = m(None, adj, logits, sp)
    assert float((out - T(g["out"])).abs().max()) < 1e-6
    assert float((van - T(g["vanilla"])).abs().max()) < 1e-6
    assert noisy.shape == out.shape and bool(torch.isfinite(noisy).all())
    # the per-library class-id table follows appended rows
    ids0 = base.class_ids().clone()
    base.add_entries(T(g["keys"])[:3], T(g["values"])[:3], torch.eye(C)[[C - 1, 0, C - 1]],
                     None if graph_level else T(g["positions"])[:3])
    ids1 = base.class_ids()
    assert ids1.numel() == ids0.numel() + 3 and ids1[-3:].tolist() == [C - 1, 0, C - 1]


def test_edge_time_encoding_host_logic(golden, cpu_ops):
    g = golden("edge_forward")
    n = int(g["X"].shape[0])
    edges, times, w = T(g["edges"]), T(g["times"]), T(g["w"])
    tn = R.relative_edge_time_encoding(edges, times, n)
    assert O.rel_err(tn, g["time_norm"]) < 1e-6
    mixed = R.relative_edge_time_encoding(edges, times, n, edge_norm=w)
    assert O.rel_err(mixed, O.edge_time_mix(w, T(g["time_norm"]))) < 1e-6
    sm = R.scatter_softmax(times.float() / 100.0, edges[:, 1], dim_size=n)
    assert O.rel_err(sm, O.scatter_softmax(times.float() / 100.0, edges[:, 1], n)) < 1e-6
    out = R.edge_rag_forward(T(g["X"]), edges, w, T(g["keys"]), T(g["values"]), int(g["num_layers"]),
                             int(g["retrieve_num"]), int(g["batch_size"]), float(g["retrieve_weight"]), edge_times=times)
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=5e-6)


def test_library_build_host_logic(golden, cpu_ops):
    g = golden("library_build")
    d, C = g["emb0"].shape[1], 3
    node = R.ToyGraphBase(None, C, d, 3, device="cpu", variant="node", capacity=4)
    graph = R.ToyGraphBase(None, C, d, 1, device="cpu", variant="graph", capacity=1, label_dtype=torch.int64)
    assert node.toy_graph_hop == int(g["node_hop"]) and graph.toy_graph_hop == int(g["graph_hop"])
    for i in range(int(g["n_graphs"])):
        node.add_graph(T(g[f"emb{i}"]), T(g[f"adj{i}"]), node_labels=T(g[f"node_labels{i}"]))
        graph.add_graph(T(g[f"emb{i}"]), T(g[f"adj{i}"]), graph_label=T(g[f"graph_label{i}"]))
    assert len(node) == g["node_keys"].shape[0] and len(graph) == 3
    assert O.rel_err(node.resource_keys, g["node_keys"]) < 1e-6 and O.rel_err(node.resource_values, g["node_values"]) < 1e-6
    assert np.array_equal(node.resource_labels.numpy(), g["node_labels"])
    assert O.rel_err(graph.resource_keys, g["graph_keys"]) < 1e-6 and O.rel_err(graph.resource_values, g["graph_values"]) < 1e-6
    assert np.array_equal(graph.resource_labels.numpy(), g["graph_labels"])
    with pytest.raises(RuntimeError, match="graph_label"):
        graph.add_graph(T(g["emb0"]), T(g["adj0"]))
    keys, values = R.make_resource_graph(T(g["edge_X"]), T(g["edge_edges"]), T(g["edge_w"]), int(g["edge_radius"]))
    np.testing.assert_allclose(keys.numpy(), g["edge_keys"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(values.numpy(), g["edge_values"], rtol=0, atol=2e-6)


def test_graph_forward_host_logic(golden, cpu_ops):
    """RAGraph(variant="graph"): one query = mean of the node embeddings, 1-hop propagation averaged over the nodes,
    weights 0.3 / 0.3 (RAGraph_graph/RAGraph.py:48-75)."""
    g = golden("graph_forward")
    d, C = g["keys"].shape[1], g["labels"].shape[1]

    class PM:
        def inference(self, features, adj):
            return T(g["emb_q"])

    base = R.ToyGraphBase(None, C, d, 1, device="cpu", variant="graph")
    assert base.retrieve_num == int(g["retrieve_num"])
    base.add_entries(T(g["keys"]), T(g["values"]), T(g["labels"]))
    model = R.RAGraph(PM(), base, 0, C, d, variant="graph").eval()
    with torch.no_grad():
        model.decoder.fc1.weight.copy_(T(g["w1"])); model.decoder.fc1.bias.copy_(T(g["b1"]))
        model.decoder.fc2.weight.copy_(T(g["w2"])); model.decoder.fc2.bias.copy_(T(g["b2"]))
        out = model(None, T(g["adj"]))
        model.finetune = False
        van = model(None, T(g["adj"]))
    assert out.shape == (1, C)
    assert float((out - T(g["logits"])).abs().max()) < 1e-6
    assert float((van - T(g["vanilla"])).abs().max()) < 1e-6


def test_edge_forward_noisy_host_logic(golden, cpu_ops):
    """use_noise and training (modules/RAGraph.py:296,310,316-322): same CPU RNG stream as the reference."""
    g = golden("edge_forward")
    torch.manual_seed(int(g["noise_seed"]))
    out = R.edge_rag_forward(T(g["X"]), T(g["edges"]), T(g["w"]), T(g["keys"]), T(g["values"]), int(g["num_layers"]),
                             int(g["retrieve_num"]), int(g["batch_size"]), float(g["retrieve_weight"]),
                             edge_times=T(g["times"]), add_noise=True, noise_retrieve_num=1)
    np.testing.assert_allclose(out.numpy(), g["out_noise"], rtol=0, atol=5e-6)


def test_inverse_sampling_host_logic(golden, cpu_ops):
    """PageRank as F = 1 SpMM steps on the transposed transition CSR; dense, sym-normalised and torch-sparse inputs."""
    g = golden("inverse_sampling")
    IS = R.InverseSampling
    adj = T(g["adj"])
    assert O.rel_err(IS.pagerank_algorithm(adj), g["pagerank"]) < 1e-5
    assert O.rel_err(IS.degree_centrality_algorithm(adj), g["degree_centrality"]) < 1e-6
    sp = IS.compute_sample_prob(adj)
    assert O.rel_err(sp, g["sample_prob"]) < 1e-5 and abs(float(sp.sum()) - 1.0) < 1e-5
    assert O.rel_err(IS.compute_sample_prob(T(g["adj_norm"])), g["sample_prob_norm"]) < 1e-5
    n = int(g["sparse_n"])
    adj_sp = torch.sparse_coo_tensor(T(g["sparse_indices"]), T(g["sparse_values"]), (n, n)).coalesce()
    assert O.rel_err(IS.pagerank_algorithm(adj_sp), g["pagerank_sparse"]) < 1e-5
    assert O.rel_err(IS.compute_sample_prob(adj_sp), g["sample_prob_sparse"]) < 1e-5


def test_fewshot_positions_derived_from_adjacency(golden, cpu_ops):
    """retrieve(search_keys, search_adj, add_noise) with the reference's three arguments: the few-shot variant derives
    the query graph's position-aware codes from search_adj with the same seeded CPU randint as the reference."""
    g = golden("fewshot_forward_node")
    from ragraph_b200.ragraph_utils import PositionAwareEncoder
    torch.manual_seed(516)                                   # the seed the golden generator used for this forward
    sp = PositionAwareEncoder.encode_position_aware_code(T(g["adj"]), 10, 10)
    assert np.array_equal(sp.numpy(), g["search_positions"])
    d, C = g["keys"].shape[1], g["labels"].shape[1]
    base = R.ToyGraphBase(None, C, d, int(g["hop"]), device="cpu", variant="node_fewshot")
    base.retrieve_num = int(g["retrieve_num"])
    base.add_entries(T(g["keys"]), T(g["values"]), T(g["labels"]), T(g["positions"]))
    torch.manual_seed(516)
    emb_a, lab_a = base.retrieve(T(g["emb_q"]), T(g["adj"]), False)                       # positions derived
    emb_b, lab_b = base.retrieve(T(g["emb_q"]), T(g["adj"]), False, T(g["search_positions"]))
    assert torch.equal(emb_a, emb_b) and torch.equal(lab_a, lab_b)
    node = R.ToyGraphBase(None, C, d, 3, device="cpu", variant="node")
    assert node.query_positions(T(g["adj"])) is None         # single-metric variants never compute them
