"""Pin the oracle against outputs of the UNMODIFIED reference (tests/golden/*.npz, made by
oracle/make_golden.py from /root/reference).  CPU only."""
import numpy as np
import torch

from oracle import ragraph_oracle as O

T = torch.from_numpy


def test_cosine_and_node_retrieve(golden):
    g = golden("node_retrieve")
    q, keys, values, labels = T(g["q"]), T(g["keys"]), T(g["values"]), T(g["labels"])
    S = O.cosine_similarity(q, keys)
    assert np.array_equal(S.numpy(), g["cosine"])            # same torch calls -> bit-identical
    assert O.rel_err(S.numpy(), O.cosine_similarity_f64(q, keys)) < 1e-6
    k = int(g["retrieve_num"])
    _, idx, emb, lab = O.retrieve(q, keys, values, labels, k)
    assert np.array_equal(emb.numpy(), g["rag_embeddings"])
    assert np.array_equal(lab.numpy(), g["rag_labels"])
    ok, bad = O.topk_sets_match(idx.numpy(), O.cosine_similarity_f64(q, keys), k)
    assert ok, bad
    _, _, emb_n, lab_n = O.retrieve(q, keys, values, labels, 2 * k, noise_indices=T(g["noise_indices"]))
    assert np.array_equal(emb_n.numpy(), g["rag_embeddings_noise"])
    assert np.array_equal(lab_n.numpy(), g["rag_labels_noise"])
    assert np.all(S.numpy()[5] == 0.0)                        # zero query row -> eps clamp, not NaN


def test_graph_retrieve_1d_query(golden):
    g = golden("graph_retrieve")
    _, idx, emb, lab = O.retrieve_graph(T(g["q"]), T(g["keys"]), T(g["values"]), T(g["labels"]),
                                        int(g["retrieve_num"]))
    assert idx.shape == (1, 3)
    assert np.array_equal(emb.numpy(), g["rag_embeddings"])
    assert np.array_equal(lab.numpy(), g["rag_labels"])


def test_fewshot_two_metric_retrieve(golden):
    g = golden("fewshot_retrieve")
    _, _, emb, lab = O.retrieve_two_metric(T(g["q"]), T(g["search_positions"]), T(g["keys"]),
                                           T(g["positions"]), T(g["values"]), T(g["labels"]),
                                           int(g["retrieve_num"]))
    assert np.array_equal(emb.numpy(), g["rag_embeddings"])
    assert np.array_equal(lab.numpy(), g["rag_labels"])


def test_propagation(golden):
    g = golden("propagation")
    for k in (0, 1, 2, 3):
        out = O.aggregate_k_hop_features(T(g["adj"]), T(g["x"]), k)
        assert np.array_equal(out.numpy(), g[f"out_k{k}"])
        assert O.rel_err(out.numpy(), O.aggregate_k_hop_features_f64(g["adj"], g["x"], k)) < 1e-6


def test_gcn_layer(golden):
    g = golden("gcn_layer")
    out = O.gcn_layer(T(g["seq"]), T(g["adj"]), T(g["weight"]), T(g["bias"]), T(g["alpha"]))
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=1e-6)


def test_edge_agg(golden):
    g = golden("edge_agg")
    Y = O.edge_agg(T(g["X"]), T(g["edges"]), T(g["w"]), int(g["num_nodes"]))
    # scatter_add_ order is implementation defined: compare at fp32 round-off, and vs fp64
    np.testing.assert_allclose(Y.numpy(), g["Y"], rtol=0, atol=2e-6)
    assert O.rel_err(g["Y"], O.edge_agg_f64(g["X"], g["edges"], g["w"], int(g["num_nodes"]))) < 1e-6
    S = O.scatter_sum(T(g["X"])[T(g["edges"])[:, 0]], T(g["edges"])[:, 1], 0, int(g["num_nodes"]))
    np.testing.assert_allclose(S.numpy(), g["scatter"], rtol=0, atol=2e-6)


def test_node_forward_fusion(golden):
    g = golden("node_forward")
    keys, values, labels = T(g["keys"]), T(g["values"]), T(g["labels"])
    emb_q, adj_q = T(g["emb_q"]), T(g["adj_q"])
    _, _, emb, lab = O.retrieve(emb_q, keys, values, labels, int(g["retrieve_num"]))
    params = (T(g["w1"]), T(g["b1"]), T(g["w2"]), T(g["b2"]))
    out = O.fuse_node(emb_q, adj_q, emb, lab, params)
    np.testing.assert_allclose(out.numpy(), g["logits"], rtol=0, atol=1e-6)
    van = O.fuse_node(emb_q, adj_q, emb, lab, params, finetune=False)
    assert np.array_equal(van.numpy(), g["vanilla"])


def test_edge_forward(golden):
    g = golden("edge_forward")
    w = T(g["w"]) * 1 / 2 + T(g["time_norm"]) * 1 / 2      # modules/RAGraph.py:267
    out = O.edge_forward(T(g["X"]), T(g["edges"]), w, T(g["keys"]), T(g["values"]),
                         int(g["num_layers"]), int(g["retrieve_num"]), int(g["batch_size"]),
                         float(g["retrieve_weight"]))
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=5e-6)


def test_merge_topk_matches_single_shard():
    g = torch.Generator().manual_seed(3)
    q = torch.randn(9, 16, generator=g); keys = torch.randn(101, 16, generator=g)
    S = O.cosine_similarity(q, keys)
    ref_s, ref_i = O.topk(S, 5)
    parts_s, parts_i = [], []
    for lo, hi in ((0, 34), (34, 68), (68, 101)):
        s, i = O.topk(S[:, lo:hi], 5)
        parts_s.append(s); parts_i.append(i + lo)
    ms, mi = O.merge_topk(torch.stack(parts_s), torch.stack(parts_i), 5)
    assert torch.equal(ms, ref_s) and torch.equal(mi, ref_i)


def test_edge_eval_ranking(golden):
    g = golden("edge_eval")
    s, i = O.rating_topk(T(g["U"]), T(g["I"]), g["hist_rowptr"], g["hist_items"], int(g["k"]))
    assert np.array_equal(i.numpy(), g["top_items"]) and np.array_equal(s.numpy(), g["top_scores"])


def test_edge_time_encoding(golden):
    g = golden("edge_forward")
    edges, times = T(g["edges"]), T(g["times"])
    tn = O.relative_edge_time_encoding(edges, times, int(g["X"].shape[0]))
    np.testing.assert_allclose(tn.numpy(), g["time_norm"], rtol=0, atol=1e-7)
    # every destination's weights sum to one
    sums = torch.zeros(int(g["X"].shape[0])).index_add_(0, edges[:, 1], tn)
    present = torch.zeros(int(g["X"].shape[0])).index_add_(0, edges[:, 1], torch.ones_like(tn)) > 0
    assert torch.allclose(sums[present], torch.ones(int(present.sum())), atol=1e-5)
    w = O.edge_time_mix(T(g["w"]), tn)
    out = O.edge_forward(T(g["X"]), edges, w, T(g["keys"]), T(g["values"]), int(g["num_layers"]),
                         int(g["retrieve_num"]), int(g["batch_size"]), float(g["retrieve_weight"]))
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=5e-6)


def _fewshot_oracle(g, graph_level):
    emb_q, adj = T(g["emb_q"]), T(g["adj"])
    keys, values, labels = T(g["keys"]), T(g["values"]), T(g["labels"])
    k = int(g["retrieve_num"])
    if graph_level:
        _, _, emb, lab = O.retrieve(emb_q, keys, values, labels, k)
    else:
        _, _, emb, lab = O.retrieve_two_metric(emb_q, T(g["search_positions"]), keys, T(g["positions"]), values, labels, k)
    dec = (T(g["dec_weight"]), T(g["dec_bias"]), T(g["dec_alpha"]))
    args = (emb_q, adj, emb, lab, T(g["mean_fewshot_logits"]), dec, float(g["retrieve_weight"]), float(g["label_weight"]),
            int(g["hop"]))
    return O.fuse_fewshot(*args, finetune=True, graph_level=graph_level), O.fuse_fewshot(*args, finetune=False,
                                                                                        graph_level=graph_level)


def test_fewshot_forward_node(golden):
    g = golden("fewshot_forward_node")
    out, van = _fewshot_oracle(g, False)
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=1e-6)
    assert np.array_equal(van.numpy(), g["vanilla"])


def test_fewshot_forward_graph(golden):
    g = golden("fewshot_forward_graph")
    out, van = _fewshot_oracle(g, True)
    assert out.shape == (1, 2)
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=1e-6)
    assert np.array_equal(van.numpy(), g["vanilla"])


def test_downprompt_node(golden):
    g = golden("downprompt_node")
    labels, seq, w = T(g["labels"]), T(g["seq"]), T(g["weight"])
    n = seq.shape[0]
    ave0 = O.average_emb(labels, T(g["feature"]).squeeze(), 3, n // 2)
    assert np.array_equal(ave0.numpy(), g["ave_init"])
    prompted = O.downstream_prompt(seq, w, elu=True)
    assert np.array_equal(prompted.numpy(), g["prompted"])
    np.testing.assert_allclose(O.prototype_scores(prompted, ave0, "softmax").numpy(), g["probs_eval"], rtol=0, atol=1e-7)
    ave1 = O.average_emb(labels, prompted, 3, n // 2)
    assert np.array_equal(ave1.numpy(), g["ave_train"])
    np.testing.assert_allclose(O.prototype_scores(prompted, ave1, "softmax").numpy(), g["probs_train"], rtol=0, atol=1e-7)
    assert np.all(np.isfinite(g["probs_eval"][7]))            # zero row: eps clamp, uniform probabilities


def test_downprompt_graph(golden):
    g = golden("downprompt_graph")
    seq, sizes = T(g["seq"]), T(g["graph_sizes"])
    gemb = O.split_and_batchify_graph_feats(O.downstream_prompt(seq, T(g["weight"]), elu=False), sizes)
    assert np.array_equal(gemb.numpy(), g["graph_emb"])
    ave = O.average_emb(T(g["graph_labels"]), gemb, 6, gemb.shape[0])
    assert np.array_equal(ave.numpy(), g["ave"])
    np.testing.assert_allclose(O.prototype_scores(gemb, ave, "log_softmax").numpy(), g["log_probs"], rtol=0, atol=1e-6)


def test_library_build(golden):
    g = golden("library_build")
    nk, nv, nl, gk, gv, gl = [], [], [], [], [], []
    for i in range(int(g["n_graphs"])):
        adj, emb = T(g[f"adj{i}"]), T(g[f"emb{i}"])
        k, v, l = O.build_library_rows_node(emb, adj, T(g[f"node_labels{i}"]), int(g["node_hop"]))
        nk.append(k); nv.append(v); nl.append(l)
        k, v, l = O.build_library_rows_graph(emb, adj, T(g[f"graph_label{i}"]), 3, int(g["graph_hop"]))
        gk.append(k); gv.append(v); gl.append(l)
    assert np.array_equal(torch.cat(nk).numpy(), g["node_keys"]) and np.array_equal(torch.cat(nv).numpy(), g["node_values"])
    assert np.array_equal(torch.cat(nl).numpy(), g["node_labels"])
    assert np.array_equal(torch.cat(gk).numpy(), g["graph_keys"]) and np.array_equal(torch.cat(gv).numpy(), g["graph_values"])
    assert np.array_equal(torch.cat(gl).numpy(), g["graph_labels"])
    keys, values = O.edge_resource_graph(T(g["edge_X"]), T(g["edge_edges"]), T(g["edge_w"]), int(g["edge_radius"]))
    np.testing.assert_allclose(keys.numpy(), g["edge_keys"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(values.numpy(), g["edge_values"], rtol=0, atol=2e-6)


def test_graph_forward_fusion(golden):
    g = golden("graph_forward")
    keys, values, labels = T(g["keys"]), T(g["values"]), T(g["labels"])
    emb_q, adj = T(g["emb_q"]), T(g["adj"])
    _, _, emb, lab = O.retrieve_graph(torch.mean(emb_q, dim=0), keys, values, labels, int(g["retrieve_num"]))
    params = (T(g["w1"]), T(g["b1"]), T(g["w2"]), T(g["b2"]))
    out = O.fuse_graph(emb_q, adj, emb, lab, params)
    assert out.shape == (1, 6)
    np.testing.assert_allclose(out.numpy(), g["logits"], rtol=0, atol=1e-6)
    van = O.fuse_graph(emb_q, adj, emb, lab, params, finetune=False)
    assert np.array_equal(van.numpy(), g["vanilla"])


def test_edge_forward_noisy_branch(golden):
    g = golden("edge_forward")
    w = O.edge_time_mix(T(g["w"]), T(g["time_norm"]))
    out = O.edge_forward(T(g["X"]), T(g["edges"]), w, T(g["keys"]), T(g["values"]), int(g["num_layers"]),
                         int(g["retrieve_num"]), int(g["batch_size"]), float(g["retrieve_weight"]),
                         noise_indices=T(g["noise_indices"]))
    np.testing.assert_allclose(out.numpy(), g["out_noise"], rtol=0, atol=5e-6)
    assert not np.allclose(g["out_noise"], g["out"], atol=1e-4)       # the branch changes the result


def test_inverse_sampling(golden):
    g = golden("inverse_sampling")
    adj = T(g["adj"])
    assert np.array_equal(O.pagerank(adj.clone()).numpy(), g["pagerank"])
    assert np.array_equal(O.degree_centrality(adj).numpy(), g["degree_centrality"])
    assert np.array_equal(O.sample_prob(adj).numpy(), g["sample_prob"])
    assert np.array_equal(O.sample_prob(T(g["adj_norm"])).numpy(), g["sample_prob_norm"])
    assert abs(float(g["sample_prob"].sum()) - 1.0) < 1e-5 and abs(float(g["pagerank"].sum()) - 1.0) < 1e-4
    # the sparse variant of the edge package computes the same quantity: dense restatement on the densified matrix
    n = int(g["sparse_n"])
    dense = torch.zeros(n, n).index_put_((T(g["sparse_indices"])[0], T(g["sparse_indices"])[1]), T(g["sparse_values"]))
    np.testing.assert_allclose(O.pagerank(dense.clone()).numpy(), g["pagerank_sparse"], rtol=0, atol=2e-7)
    np.testing.assert_allclose(O.sample_prob(dense).numpy(), g["sample_prob_sparse"], rtol=0, atol=2e-7)
