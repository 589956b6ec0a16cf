"""GPU parity tests for the rows widened after the core path: downstream prompt (section 8 a9), few-shot fusion (a7,
few-shot variants), edge time encoding (K10).  Same bar as tests/test_gpu_parity.py: everything goes through
torch.ops.ragraph -> C ABI and is compared with the oracle / the reference-generated golden vectors."""
import numpy as np
import pytest
import torch

import ragraph_b200 as R
from ragraph_b200 import _lib as L
from ragraph_b200 import downprompt as DP
from ragraph_b200 import ops
from oracle import ragraph_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
T = torch.from_numpy
REL = 1e-5


def cu(a):
    return (T(a) if isinstance(a, np.ndarray) else a).to(DEV)


# ------------------------------------------------------------------------------------------ K10 time encoding
def test_scatter_softmax_matches_oracle():
    g = torch.Generator().manual_seed(5)
    E, n = 20000, 700
    index = torch.randint(0, n, (E,), generator=g)
    index[index == 13] = 14                                   # an empty group
    src = torch.randn(E, generator=g) * 3
    out = ops.scatter_softmax(cu(src), cu(index), n).cpu()
    ref = O.scatter_softmax(src, index, n)
    assert O.rel_err(out, ref) < REL
    sums = torch.zeros(n).index_add_(0, index, out)
    present = torch.bincount(index, minlength=n) > 0
    assert torch.allclose(sums[present], torch.ones(int(present.sum())), atol=1e-5)


def test_scatter_softmax_edge_cases():
    assert ops.scatter_softmax(torch.empty(0, device=DEV), torch.empty(0, dtype=torch.int64, device=DEV), 5).numel() == 0
    one = ops.scatter_softmax(torch.tensor([3.0, -2.0, 100.0], device=DEV), torch.tensor([0, 1, 2], device=DEV), 3)
    assert torch.equal(one.cpu(), torch.ones(3))              # singleton groups, large magnitudes: no overflow
    big = ops.scatter_softmax(torch.tensor([1000.0, 999.0], device=DEV), torch.tensor([0, 0], device=DEV), 1).cpu()
    assert O.rel_err(big, torch.softmax(torch.tensor([1000.0, 999.0]), 0)) < REL
    with pytest.raises(RuntimeError):
        ops.scatter_softmax(torch.zeros(3, device=DEV), torch.zeros(2, dtype=torch.int64, device=DEV), 1)


def test_edge_time_encoding_golden(golden):
    g = golden("edge_forward")
    n = int(g["X"].shape[0])
    tn = R.relative_edge_time_encoding(cu(g["edges"]), cu(g["times"]), n).cpu()
    assert O.rel_err(tn, g["time_norm"]) < REL
    mixed = R.relative_edge_time_encoding(cu(g["edges"]), cu(g["times"]), n, edge_norm=cu(g["w"])).cpu()
    assert O.rel_err(mixed, O.edge_time_mix(T(g["w"]), T(g["time_norm"]))) < REL
    # max_step given (RAGraph.forward(..., max_time_step)): oracle restatement
    ms = float(g["times"].max()) * 1.5
    tn2 = R.relative_edge_time_encoding(cu(g["edges"]), cu(g["times"]), n, max_step=ms).cpu()
    assert O.rel_err(tn2, O.relative_edge_time_encoding(T(g["edges"]), T(g["times"]), n, torch.tensor(ms))) < REL


def test_edge_forward_with_times_golden(golden):
    """The whole modules/RAGraph.py:265-333 forward from raw edge times: time softmax + mix, 3 LightGCN layers,
    per-batch retrieve + mean + blend, against the unmodified reference's output."""
    g = golden("edge_forward")
    out = R.edge_rag_forward(cu(g["X"]), cu(g["edges"]), cu(g["w"]), cu(g["keys"]), cu(g["values"]),
                             int(g["num_layers"]), int(g["retrieve_num"]), int(g["batch_size"]),
                             float(g["retrieve_weight"]), edge_times=cu(g["times"])).cpu()
    assert O.rel_err(out, g["out"]) < REL


# ------------------------------------------------------------------------------------------ a9 downstream prompt
@pytest.mark.parametrize("d", [24, 64, 250, 256])
@pytest.mark.parametrize("act", [L.ACT_NONE, L.ACT_ELU])
def test_prompt_act(d, act):
    g = torch.Generator().manual_seed(d + act)
    x = torch.randn(333, d, generator=g) * 2
    w = torch.randn(1, d, generator=g)
    out = ops.prompt_act(cu(x), cu(w), act).cpu()
    ref = O.downstream_prompt(x, w, elu=bool(act))
    assert out.shape == ref.shape and O.rel_err(out, ref) < 1e-6
    if act == L.ACT_NONE:
        assert torch.equal(out, ref)                          # one fp32 multiply: bit exact


@pytest.mark.parametrize("C,d,n", [(2, 24, 50), (3, 256, 777), (6, 64, 200), (7, 250, 129), (20, 32, 64), (32, 16, 40)])
@pytest.mark.parametrize("mode", ["raw", "softmax", "log_softmax"])
def test_prototype_scores(C, d, n, mode):
    g = torch.Generator().manual_seed(C * 1000 + d)
    x = torch.randn(n, d, generator=g)
    x[3] = 0.0                                                # eps clamp row
    proto = torch.randn(C, d, generator=g)
    m = {"raw": L.SCORES_RAW, "softmax": L.SCORES_SOFTMAX, "log_softmax": L.SCORES_LOG_SOFTMAX}[mode]
    out = ops.prototype_scores(cu(x), cu(proto), m).cpu()
    xs = x[:64] if n > 64 else x                              # the oracle is a Python double loop
    ref = O.prototype_scores(xs, proto, mode)
    assert out.shape == (n, C)
    assert float((out[:xs.shape[0]] - ref).abs().max()) < 2e-6
    if mode == "softmax":
        assert torch.allclose(out.sum(1), torch.ones(n), atol=1e-5)
    # prompt applied on the fly == prompt materialised first
    w = torch.randn(d, generator=g)
    fused = ops.prototype_scores(cu(x), cu(proto), m, cu(w), L.ACT_ELU).cpu()
    two = ops.prototype_scores(ops.prompt_act(cu(x), cu(w), L.ACT_ELU), cu(proto), m).cpu()
    assert float((fused - two).abs().max()) < 2e-6


def test_prototype_scores_limits():
    x = torch.randn(4, 8, device=DEV)
    with pytest.raises(RuntimeError, match="classes"):
        ops.prototype_scores(x, torch.randn(33, 8, device=DEV))
    with pytest.raises(RuntimeError):
        ops.prototype_scores(x, torch.randn(3, 9, device=DEV))
    assert ops.prototype_scores(torch.empty(0, 8, device=DEV), torch.randn(3, 8, device=DEV)).shape == (0, 3)


def test_downprompt_node_golden(golden):
    g = golden("downprompt_node")
    d = g["seq"].shape[1]
    p = torch.zeros(1, d, device=DEV)
    m = DP.downprompt(p, p, p, d, 3, cu(g["feature"]), cu(g["labels"])).to(DEV)
    m.downprompt.weight.data.copy_(cu(g["weight"]))
    assert O.rel_err(m.ave.cpu(), g["ave_init"]) < REL
    assert O.rel_err(m.downprompt(cu(g["seq"])).cpu(), g["prompted"]) < 1e-6
    probs = m(cu(g["seq"]), 0).cpu()
    assert float((probs - T(g["probs_eval"])).abs().max()) < 2e-6
    probs_t = m(cu(g["seq"]), 1).cpu()
    assert O.rel_err(m.ave.cpu(), g["ave_train"]) < REL
    assert float((probs_t - T(g["probs_train"])).abs().max()) < 2e-6


def test_downprompt_graph_golden(golden):
    g = golden("downprompt_graph")
    d = g["seq"].shape[1]
    p = torch.zeros(1, d, device=DEV)
    m = DP.downprompt(p, p, p, d, 6).to(DEV)
    m.downprompt.weight.data.copy_(cu(g["weight"]))
    gemb = m(cu(g["seq"]), cu(g["graph_sizes"]))
    assert O.rel_err(gemb.cpu(), g["graph_emb"]) < REL
    ave = DP.averageemb(cu(g["graph_labels"]), gemb, 6, slots=gemb.shape[0])
    assert O.rel_err(ave.cpu(), g["ave"]) < REL
    logp = DP.predict(gemb.shape[0], 6, gemb, ave).cpu()
    assert float((logp - T(g["log_probs"])).abs().max()) < 5e-6


def test_split_and_batchify_ragged():
    g = torch.Generator().manual_seed(77)
    sizes = torch.tensor([1, 0, 40, 3, 1500, 2])                # an empty graph, a long one (split-row path)
    x = torch.randn(int(sizes.sum()), 48, generator=g)
    out = DP.split_and_batchify_graph_feats(cu(x), cu(sizes)).cpu()
    ref = O.split_and_batchify_graph_feats(x, sizes)
    assert O.rel_err(out, ref) < REL and torch.all(out[1] == 0)


# ------------------------------------------------------------------------------------------ a7 few-shot fusion
class _FewShotBackbone:
    """encode returns the fixed query embeddings; decode is the second GCN layer through the CUDA GCN mirror."""

    def __init__(self, g):
        self.emb = cu(g["emb_q"])
        d, C = g["dec_weight"].shape[1], g["dec_weight"].shape[0]
        self.layer = R.GCN(d, C, "prelu").to(DEV)
        with torch.no_grad():
            self.layer.fc.weight.copy_(cu(g["dec_weight"]))
            self.layer.bias.copy_(cu(g["dec_bias"]))
            self.layer.act.weight.copy_(cu(g["dec_alpha"]))

    def encode(self, features, adj):
        return self.emb

    def decode(self, hidden, adj):
        return self.layer((hidden, adj))


def _fewshot_model(g, graph_level):
    d, C = g["keys"].shape[1], g["labels"].shape[1]
    base = R.ToyGraphBase(None, C, d, int(g["hop"]), device=DEV, variant="node" if graph_level else "node_fewshot")
    base.retrieve_num = int(g["retrieve_num"])
    base.add_entries(cu(g["keys"]), cu(g["values"]), cu(g["labels"]), None if graph_level else cu(g["positions"]))
    return R.RAGraphFewShot(_FewShotBackbone(g), base, d, finetune=True, query_graph_hop=int(g["hop"]),
                            retrieve_weight=float(g["retrieve_weight"]), label_weight=float(g["label_weight"]),
                            graph_level=graph_level).eval()


@pytest.mark.parametrize("graph_level", [False, True])
def test_fewshot_forward_golden(golden, graph_level):
    g = golden("fewshot_forward_graph" if graph_level else "fewshot_forward_node")
    m = _fewshot_model(g, graph_level)
    sp = None if graph_level else cu(g["search_positions"])
    with torch.no_grad():
        out = m(None, cu(g["adj"]), cu(g["mean_fewshot_logits"]), sp).cpu()
        m.finetune = False
        van = m(None, cu(g["adj"]), cu(g["mean_fewshot_logits"]), sp).cpu()
    assert out.shape == g["out"].shape and van.shape == g["vanilla"].shape
    assert float((out - T(g["out"])).abs().max()) < 5e-6
    assert float((van - T(g["vanilla"])).abs().max()) < 1e-6


def test_fewshot_noisy_branch_shapes(golden):
    g = golden("fewshot_forward_node")
    m = _fewshot_model(g, False)
    m.noise_finetune = True
    m.train()
    with torch.no_grad():
        out = m(None, cu(g["adj"]), cu(g["mean_fewshot_logits"]), cu(g["search_positions"]))
    assert out.shape == g["out"].shape and bool(torch.isfinite(out).all())


# ------------------------------------------------------------------------------------------ a8 / 8f-2 library construction
@pytest.mark.parametrize("d", [24, 64, 250])
def test_rows_normalize(d):
    g = torch.Generator().manual_seed(d)
    x = torch.randn(500, d, generator=g) * 3
    x[7] = 0.0                                                # zero row: eps clamp, stays zero
    out = ops.rows_normalize(cu(x)).cpu()
    ref = torch.nn.functional.normalize(x, p=2, dim=-1)
    assert O.rel_err(out, ref) < 1e-6 and bool((out[7] == 0).all())


def test_library_build_golden(golden):
    g = golden("library_build")
    d, C = g["emb0"].shape[1], 3
    node = R.ToyGraphBase(None, C, d, 3, device=DEV, variant="node", capacity=4)
    graph = R.ToyGraphBase(None, C, d, 1, device=DEV, variant="graph", capacity=1, label_dtype=torch.int64)
    for i in range(int(g["n_graphs"])):
        node.add_graph(cu(g[f"emb{i}"]), cu(g[f"adj{i}"]), node_labels=cu(g[f"node_labels{i}"]))
        graph.add_graph(cu(g[f"emb{i}"]), cu(g[f"adj{i}"]), graph_label=cu(g[f"graph_label{i}"]))
    assert O.rel_err(node.resource_keys.cpu(), g["node_keys"]) < REL
    assert O.rel_err(node.resource_values.cpu(), g["node_values"]) < REL
    assert np.array_equal(node.resource_labels.cpu().numpy(), g["node_labels"])
    assert O.rel_err(graph.resource_keys.cpu(), g["graph_keys"]) < REL
    assert O.rel_err(graph.resource_values.cpu(), g["graph_values"]) < REL
    assert np.array_equal(graph.resource_labels.cpu().numpy(), g["graph_labels"])
    keys, values = R.make_resource_graph(cu(g["edge_X"]), cu(g["edge_edges"]), cu(g["edge_w"]), int(g["edge_radius"]))
    assert O.rel_err(keys.cpu(), g["edge_keys"]) < REL and O.rel_err(values.cpu(), g["edge_values"]) < REL
    # a library built this way answers queries like the reference's: its own rows are their own nearest neighbours
    _, idx = node.topk(node.resource_keys[:20].clone(), 1)
    assert torch.equal(idx.cpu().reshape(-1), torch.arange(20))


# ------------------------------------------------------------------------------------------ a7 graph-variant forward
def test_graph_forward_golden(golden):
    g = golden("graph_forward")
    d, C = g["keys"].shape[1], g["labels"].shape[1]

    class PM:
        def inference(self, features, adj):
            return cu(g["emb_q"])

    base = R.ToyGraphBase(None, C, d, 1, device=DEV, variant="graph")
    base.add_entries(cu(g["keys"]), cu(g["values"]), cu(g["labels"]))
    model = R.RAGraph(PM(), base, 0, C, d, variant="graph").to(DEV).eval()
    with torch.no_grad():
        model.decoder.fc1.weight.copy_(cu(g["w1"])); model.decoder.fc1.bias.copy_(cu(g["b1"]))
        model.decoder.fc2.weight.copy_(cu(g["w2"])); model.decoder.fc2.bias.copy_(cu(g["b2"]))
        out = model(None, cu(g["adj"])).cpu()
        model.finetune = False
        van = model(None, cu(g["adj"])).cpu()
    assert out.shape == (1, C)
    assert float((out - T(g["logits"])).abs().max()) < 1e-5
    assert float((van - T(g["vanilla"])).abs().max()) < 1e-6


def test_edge_forward_noisy_golden(golden):
    g = golden("edge_forward")
    torch.manual_seed(int(g["noise_seed"]))                     # the noise rows come from the CPU generator, like the reference
    out = R.edge_rag_forward(cu(g["X"]), cu(g["edges"]), cu(g["w"]), cu(g["keys"]), cu(g["values"]),
                             int(g["num_layers"]), int(g["retrieve_num"]), int(g["batch_size"]),
                             float(g["retrieve_weight"]), edge_times=cu(g["times"]), add_noise=True).cpu()
    assert O.rel_err(out, g["out_noise"]) < REL


# ------------------------------------------------------------------------------------------ 8f-2 PageRank as SpMV
def test_inverse_sampling_golden(golden):
    # The power iteration stops when the L1 change drops below 1e-6: a different fp32 summation order can move the stop
    # by one step, which moves an entry by up to ~1e-6 / n of absolute value -- hence 5e-5 relative to the largest entry.
    tol = 5e-5
    g = golden("inverse_sampling")
    IS = R.InverseSampling
    assert O.rel_err(IS.pagerank_algorithm(cu(g["adj"])).cpu(), g["pagerank"]) < tol
    sp = IS.compute_sample_prob(cu(g["adj"])).cpu()
    assert O.rel_err(sp, g["sample_prob"]) < tol and abs(float(sp.sum()) - 1.0) < 1e-5
    assert O.rel_err(IS.compute_sample_prob(cu(g["adj_norm"])).cpu(), g["sample_prob_norm"]) < tol
    n = int(g["sparse_n"])
    adj_sp = torch.sparse_coo_tensor(cu(g["sparse_indices"]), cu(g["sparse_values"]), (n, n)).coalesce()
    assert O.rel_err(IS.compute_sample_prob(adj_sp).cpu(), g["sample_prob_sparse"]) < tol


# ------------------------------------------------------------------------------------------ few-shot helpers
def test_fewshot_helpers_vs_oracle():
    g = torch.Generator().manual_seed(123)
    C, Ld, n_support, n = 3, 12, 15, 300
    support_logits = torch.randn(n_support, Ld, generator=g)
    support_labels = (torch.arange(n_support) % C)[torch.randperm(n_support, generator=g)]
    logits = torch.randn(n, Ld, generator=g)
    F_ = R.fewshot
    means, uniq = F_.fewshot_mean(cu(support_logits), cu(support_labels))
    ref_means, ref_uniq = O.fewshot_mean(support_logits, support_labels)
    assert torch.equal(uniq.cpu(), ref_uniq) and O.rel_err(means.cpu(), ref_means) < REL
    table = F_.fewshot_mean_logits(cu(support_logits), cu(support_labels))
    sim = F_.fewshot_predict_logits(table, cu(logits)).cpu()
    ref_sim = O.fewshot_predict_logits(ref_means, logits)
    assert float((sim - ref_sim).abs().max()) < 2e-6
    pred = F_.fewshot_predict_labels_by_mean(table, cu(logits)).cpu()
    margin = ref_sim.topk(2, dim=1).values
    clear = (margin[:, 0] - margin[:, 1]) > 1e-5                # rows whose best class is not a near tie
    assert torch.equal(pred[clear], ref_sim.argmax(1)[clear])
    loss = F_.fewshot_predict_loss(cu(support_logits), cu(support_labels), cu(logits), cu(torch.randint(0, C, (n,), generator=g)))
    assert bool(torch.isfinite(loss))


# ---- 8f rank 3: host preprocessing on the device ------------------------------------------------------------------------
def test_negative_sampling_distribution_and_exclusion():
    """Device negative sampler vs the reference's rejection loop (RAGraph_edge/utils/dataloader.py:140-152): never an item
    of the user's history, uniform over the rest (chi-square against the exact complement distribution, and against an
    oracle run of the reference loop with numpy's generator), deterministic per seed, user-major order."""
    import numpy as np
    from ragraph_b200 import edge as E
    rng = np.random.default_rng(0)
    U, I = 50, 200
    hist = {u: sorted(rng.choice(I, size=int(rng.integers(0, 150)), replace=False).tolist()) for u in range(U)}
    hist[7] = list(range(I - 1))                                     # a user with ONE admissible item
    rowptr, items = E.history_csr(hist, U, DEV)
    users = torch.arange(U, device=DEV).repeat_interleave(400)
    out = E.negative_sampling(users, rowptr, items, I, n=3, seed=123).reshape(-1, 3).cpu().numpy()
    assert out.shape == (U * 400, 3) and out.min() >= 0 and out.max() < I
    again = E.negative_sampling(users, rowptr, items, I, n=3, seed=123).reshape(-1, 3).cpu().numpy()
    assert np.array_equal(out, again)
    other = E.negative_sampling(users, rowptr, items, I, n=3, seed=124).reshape(-1, 3).cpu().numpy()
    assert not np.array_equal(out, other)
    uu = users.cpu().numpy()
    for u in range(U):
        got = out[uu == u].reshape(-1)
        assert not set(got.tolist()) & set(hist[u]), u
        allowed = np.setdiff1d(np.arange(I), np.array(hist[u], dtype=np.int64))
        cnt = np.array([(got == a).sum() for a in allowed], dtype=np.float64)
        exp = got.size / allowed.size
        if allowed.size > 1:
            chi2 = ((cnt - exp) ** 2 / exp).sum()
            assert chi2 < allowed.size + 6 * np.sqrt(2 * allowed.size) + 10, (u, chi2, allowed.size)   # ~6 sigma
    assert set(out[uu == 7].reshape(-1).tolist()) == {I - 1}
    # the reference loop itself (numpy generator) on one user: same support, same uniformity
    np.random.seed(0)
    ref = []
    for _ in range(1200):
        while True:
            neg = np.random.randint(low=0, high=I, size=1)[0]
            if neg not in set(hist[3]):
                break
        ref.append(neg)
    assert set(ref) <= set(np.setdiff1d(np.arange(I), hist[3]).tolist())


def test_process_graph_batch_on_device_matches_reference_dense():
    """process_graph_batch with CUDA inputs: CSR of D^-1/2 (A + I) D^-1/2 == the reference's dense block-diagonal matrix
    (oracle restatement of process_tu_dataset, RAGraph_node/ragraph_utils/utility.py:30-72), and propagating over it equals
    propagating over the dense adjacency."""
    import numpy as np
    from ragraph_b200 import process_graph_batch
    rng = np.random.default_rng(1)
    xs, eis = [], []
    for n in (7, 12, 1, 30):
        xs.append(torch.tensor(rng.random((n, 9)), dtype=torch.float32))
        m = max(1, 2 * n)
        e = rng.integers(0, n, size=(2, m))
        eis.append(torch.tensor(np.concatenate([e, e[::-1]], axis=1)))           # symmetric, duplicates allowed
    f0, adj0, l0 = O.process_tu_arrays([x.numpy() for x in xs], [e.numpy() for e in eis], 6)
    feats, csr, labs = process_graph_batch([x.to(DEV) for x in xs], [e.to(DEV) for e in eis], 6)
    assert feats.is_cuda and csr.rowptr.is_cuda
    assert torch.equal(feats.cpu(), f0) and torch.equal(labs.cpu(), l0)
    assert float((csr.to_dense().cpu() - adj0).abs().max()) < 1e-6
    x = torch.randn(adj0.shape[0], 32)
    got = R.Propagation.aggregate_k_hop_features(csr, x.to(DEV), 2).cpu()
    ref = O.aggregate_k_hop_features(adj0, x, 2)
    assert O.rel_err(got.numpy(), ref.numpy()) < 1e-5


def test_fewshot_finetune_step_backpropagates_into_the_encoder(golden):
    """The few-shot fine-tune optimises the backbone too (encode() = convs[0] is trainable, Adam over rag_model.parameters(),
    RAGraph_node_fewshot/RAGraph.py:41-69): loss.backward() must reach the encoder through propagate -> blend -> decode.
    The blend with the retrieved rows runs in torch when the encoder output requires grad (gather_reduce has no autograd
    formula and the library rows carry none); gradient == the all-torch restatement of the same forward."""
    g = golden("fewshot_forward_node")
    m = _fewshot_model(g, False).train()
    emb = torch.nn.Parameter(cu(g["emb_q"]).clone())            # stands for the trainable first GCN layer's output
    m.pretrain_model.emb = emb
    adj, logits, sp = cu(g["adj"]), cu(g["mean_fewshot_logits"]), cu(g["search_positions"])
    out = m(None, adj, logits, sp)
    loss = (out * torch.linspace(-1, 1, out.numel(), device=DEV).reshape(out.shape)).sum()
    loss.backward()
    assert emb.grad is not None and bool(torch.isfinite(emb.grad).all()) and float(emb.grad.abs().sum()) > 0
    assert m.pretrain_model.layer.fc.weight.grad is not None
    # all-torch restatement (dense adjacency, advanced indexing) of the same forward for the gradient
    emb2 = cu(g["emb_q"]).clone().requires_grad_(True)
    base = m.toy_graph_base
    with torch.no_grad():
        _, idx = base.topk(emb2.detach(), base.retrieve_num, sp)
    a = adj / adj.sum(dim=1, keepdim=True)
    qe = emb2
    for _ in range(m.query_graph_hop):
        qe = torch.relu(a @ qe)
    hidden = qe * (1 - m.retrieve_weight) + base.resource_values[idx].sum(1) * m.retrieve_weight
    layer = m.pretrain_model.layer
    dec = torch.nn.functional.prelu(adj @ (hidden @ layer.fc.weight.t()) + layer.bias, layer.act.weight)
    rag_logits = logits[torch.argmax(base.resource_labels[idx], dim=-1)].mean(1)
    out2 = dec * (1 - m.label_weight) + rag_logits * m.label_weight
    assert float((out2 - out).abs().max()) < 1e-5
    (out2 * torch.linspace(-1, 1, out.numel(), device=DEV).reshape(out.shape)).sum().backward()
    assert float((emb2.grad - emb.grad).abs().max()) < 1e-4 * float(emb2.grad.abs().max())


# ---- f4 on the tensor cores: masked dot-product ranking -------------------------------------------------------------------
@pytest.mark.parametrize("B,I,d,k,heavy", [(300, 20000, 64, 20, False), (64, 9000, 64, 50, False), (200, 30000, 128, 20, True),
                                            (1100, 150000, 64, 20, True)])
def test_rating_topk_tensor_cores_equals_fp32_path(B, I, d, k, heavy):
    """rating_topk on the tensor cores (fp16 filter over the scaled item table, refine drops the history items, certificate,
    second pass / fp32 kernel for the rest) == the fp32 CUDA-core kernel == the reference's formulation (dense rating,
    history set to -inf, torch.topk; RAGraph_edge/utils/metrics.py:48-53,96-118), ties within 1e-6 aside.  ``heavy``: some
    users' histories ARE their best-scoring items (hundreds of them), so their candidate lists hold too few admissible items
    and the row must go through the second pass or the fp32 kernel; one user has a single admissible item."""
    import numpy as np
    from ragraph_b200 import edge as E
    g = torch.Generator().manual_seed(B + I)
    users = torch.randn(B, d, generator=g) * 0.3
    items = torch.randn(I, d, generator=g) * torch.rand(I, 1, generator=g) * 2.0      # un-normalised, norms vary 0..~2 sqrt(d)
    users[5] = 0.0                                                                     # a zero user: every score ties at 0
    rating = users.double() @ items.double().T
    hist = {}
    for u in range(B):
        n_h = int(torch.randint(0, 60, (1,), generator=g))
        hist[u] = torch.randperm(I, generator=g)[:n_h].tolist()
    if heavy:
        for u in range(0, B, 7):                                                       # history = the user's own top items
            hist[u] = rating[u].topk(int(torch.randint(100, 700, (1,), generator=g))).indices.tolist()
        hist[3] = [i for i in range(I) if i != 777]                                    # ONE admissible item
    rowptr, hitems = E.history_csr(hist, B, DEV)
    ud, idv = users.to(DEV), items.to(DEV)
    got_tc = E.rating_topk(ud, idv, k, rowptr, hitems, tensor_cores=True).cpu()
    got_f32 = E.rating_topk(ud, idv, k, rowptr, hitems, tensor_cores=False).cpu()
    masked = rating.clone()
    for u, h in hist.items():
        if h:
            masked[u, torch.tensor(h)] = -float("inf")
    for got in (got_tc, got_f32):
        for u in range(B):
            row = got[u]
            valid = row[row >= 0]
            n_adm = I - len(set(hist[u]))
            assert valid.numel() == min(k, n_adm), (u, valid.numel(), n_adm)
            assert not set(valid.tolist()) & set(hist[u]), u
            assert len(set(valid.tolist())) == valid.numel(), u
            if u == 5:
                continue                                                               # all ties: any admissible set is right
            ref_scores = masked[u].topk(valid.numel()).values
            got_scores = rating[u, valid]
            assert float((got_scores.sort(descending=True).values - ref_scores).abs().max()) < 1e-5 * max(1.0, float(ref_scores.abs().max())), u
    same = (got_tc == got_f32).all(dim=1).float().mean()
    assert float(same) > 0.99, float(same)
