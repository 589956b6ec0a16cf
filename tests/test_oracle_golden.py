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
