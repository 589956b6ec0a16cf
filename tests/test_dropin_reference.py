"""Drop-in check against the UNMODIFIED reference tree (only where /root/reference exists: the build container; skipped
on the GPU box): the monkey-patch recipe of INTEGRATION.md section 3 is applied to the reference's own modules and the
reference's own ``RAGraph.forward`` / ``ToyGraphBase.retrieve`` / ``_agg`` are run before and after.  The CUDA ops are
replaced by the oracle-backed stand-ins (no GPU here), so what is verified is the BOUNDARY: our mirrors accept the
reference's call sites unchanged -- names, argument order, shapes, dtypes -- and give the same results."""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

import ragraph_b200 as R
from oracle import ragraph_oracle as O
from test_host_widen import cpu_ops  # noqa: F401  (fixture)

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
_VARIANT_PKGS = ("ragraph_utils", "layers", "models", "utils", "modules", "RAGraph", "preprompt", "downprompt", "aug", "utility")


@pytest.fixture
def reference(monkeypatch):
    """Makes one reference variant importable on the CPU (same means as oracle/make_golden.py) and cleans up after."""
    def purge():
        for m in list(sys.modules):
            if m.split(".")[0] in _VARIANT_PKGS:
                del sys.modules[m]

    tg = types.ModuleType("torch_geometric")
    tgl = types.ModuleType("torch_geometric.loader"); tgl.DataLoader = object
    tgd = types.ModuleType("torch_geometric.datasets"); tgd.TUDataset = object
    tg.loader, tg.datasets = tgl, tgd
    for name, mod in (("torch_geometric", tg), ("torch_geometric.loader", tgl), ("torch_geometric.datasets", tgd)):
        monkeypatch.setitem(sys.modules, name, mod)
    ts = types.ModuleType("torch_scatter"); ts.scatter_softmax = lambda src, index, dim_size=None: O.scatter_softmax(src, index, dim_size)
    monkeypatch.setitem(sys.modules, "torch_scatter", ts)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)

    def enter(variant, argv=None):
        purge()
        monkeypatch.syspath_prepend(os.path.join(REF, variant))
        if argv is not None:
            monkeypatch.setattr(sys, "argv", argv)

    yield enter
    purge()


def _sym_norm_adj(n, p, gen):
    a = (torch.rand(n, n, generator=gen) < p).float()
    a = torch.triu(a, 1); a = a + a.t() + torch.eye(n)
    dinv = a.sum(1).pow(-0.5)
    return dinv[:, None] * a * dinv[None, :]


def test_node_forward_with_integration_patches(reference, cpu_ops):  # noqa: F811
    reference("RAGraph_node")
    sys.modules.setdefault("utils", types.ModuleType("utils")).process = None
    from ragraph_utils import Propagation as RefPropagation, TaskDecoder as RefDecoder
    ref_tgb = sys.modules["ragraph_utils.ToyGraphBase"]       # the MODULE (the package re-exports the class under that name)
    import layers.gcn as ref_gcn
    RefRAG = importlib.import_module("RAGraph").RAGraph

    g = torch.Generator().manual_seed(31)
    n, N, d, C, F_in = 35, 500, 32, 3, 20
    adj = _sym_norm_adj(n, 0.1, g)
    feats = torch.randn(n, F_in, generator=g)
    keys = torch.nn.functional.normalize(torch.randn(N, d, generator=g), dim=-1)
    values = torch.randn(N, d, generator=g)
    labels = torch.nn.functional.one_hot(torch.randint(0, C, (N,), generator=g), C).float()
    torch.manual_seed(5)
    layer = ref_gcn.GCN(F_in, d, "prelu")                       # the reference's own GCN layer as the backbone

    class PM:
        def inference(self, features, a):
            with torch.no_grad():
                return layer((features, a))

    base = ref_tgb.ToyGraphBase(None, C, d, 3)
    base.resource_keys, base.resource_values, base.resource_labels = keys, values, labels
    torch.manual_seed(6)
    shim = object.__new__(RefRAG)
    torch.nn.Module.__init__(shim)
    shim.pretrain_model, shim.toy_graph_base, shim.decoder = PM(), base, RefDecoder(d, d, C)
    shim.retrieve_weight, shim.label_weight, shim.finetune = 0.5, 0.5, True
    shim.noise_finetune, shim.query_graph_hop = False, 3
    shim.eval()
    with torch.no_grad():
        want = shim.forward(feats, adj)
        want_emb, want_lab = base.retrieve(PM().inference(feats, adj), adj, False)

    # ---- INTEGRATION.md section 3, monkey-patch form ----------------------------------------------------------
    ref_tgb.SimilarityFunctions = R.SimilarityFunctions
    RefPropagation.aggregate_k_hop_features = staticmethod(R.Propagation.aggregate_k_hop_features)
    ref_gcn.GCN.forward = R.GCN.forward
    with torch.no_grad():
        got = shim.forward(feats, adj)
        got_emb, got_lab = base.retrieve(PM().inference(feats, adj), adj, False)
    assert got.shape == want.shape and float((got - want).abs().max()) < 1e-5
    assert torch.equal(got_lab, want_lab) and float((got_emb - want_emb).abs().max()) == 0.0

    # ---- the reference's state_dict loads into our mirrors (same parameter names) -------------------------------
    ours = R.GCN(F_in, d, "prelu"); ours.load_state_dict(layer.state_dict())
    dec = R.TaskDecoder(d, d, C); dec.load_state_dict(shim.decoder.state_dict())
    with torch.no_grad():
        assert float((ours((feats, adj)) - PM().inference(feats, adj)).abs().max()) < 1e-5

    # ---- our store behind the reference's forward: ToyGraphBase swapped as a whole ------------------------------
    mine = R.ToyGraphBase(None, C, d, 3, device="cpu")
    mine.add_entries(keys, values, labels)
    shim.toy_graph_base = mine
    with torch.no_grad():
        got2 = shim.forward(feats, adj)
    assert float((got2 - want).abs().max()) < 1e-5


def test_edge_agg_and_scatter_with_integration_patches(reference, cpu_ops):  # noqa: F811
    reference("RAGraph_edge", argv=["x", "--device", "cpu", "--data_path", "dataset/amazon"])
    from modules.RAGraph import RAGraph as RefEdge
    import modules.RAGraph as ref_mod
    g = torch.Generator().manual_seed(77)
    nu, ni, d, E = 30, 25, 16, 300
    n = nu + ni
    u = torch.randint(0, nu, (E,), generator=g); i = torch.randint(0, ni, (E,), generator=g) + nu
    edges = torch.cat([torch.stack([u, i], 1), torch.stack([i, u], 1)], 0)
    w = torch.rand(edges.shape[0], generator=g)
    X = torch.randn(n, d, generator=g)
    shim = types.SimpleNamespace(num_users=nu, num_items=ni)
    want = RefEdge._agg(shim, X, edges, w)
    # `from modules.utils import scatter_sum` swapped for ours inside the reference module: _agg is untouched
    ref_mod.scatter_sum = R.scatter_sum
    got = RefEdge._agg(shim, X, edges, w)
    assert float((got - want).abs().max()) < 5e-6
    # or the aggregator object in place of the method
    agg = R.EdgeAggregator(n)
    assert float((agg(X, edges, w) - want).abs().max()) < 5e-6
    times = torch.randint(0, 1000, (edges.shape[0],), generator=g)
    want_t = RefEdge._relative_edge_time_encoding(shim, edges, times)
    assert float((R.relative_edge_time_encoding(edges, times, n) - want_t).abs().max()) < 1e-6


def test_graph_forward_with_our_store(reference, cpu_ops):  # noqa: F811
    """RAGraph_graph/RAGraph.py:48-75 unchanged, our ToyGraphBase (graph variant: 1-D query, int64 one-hot labels) behind it."""
    reference("RAGraph_graph")
    from ragraph_utils import TaskDecoder as RefDecoder
    ref_tgb = sys.modules["ragraph_utils.ToyGraphBase"]
    RefRAG = importlib.import_module("RAGraph").RAGraph
    g = torch.Generator().manual_seed(8)
    n, N, d, C = 19, 200, 32, 6
    adj = _sym_norm_adj(n, 0.2, g)
    emb = torch.randn(n, d, generator=g)
    keys = torch.randn(N, d, generator=g) * 0.3
    values = torch.randn(N, d, generator=g)
    labels = torch.nn.functional.one_hot(torch.randint(0, C, (N,), generator=g), C)

    class PM:
        def inference(self, features, a):
            return emb

    base = ref_tgb.ToyGraphBase(None, C, d, 1)
    base.resource_keys, base.resource_values = keys, values
    base.resource_labels = torch.cat((torch.empty(0, C), labels), dim=0)
    torch.manual_seed(9)
    shim = object.__new__(RefRAG)
    torch.nn.Module.__init__(shim)
    shim.pretrain_model, shim.toy_graph_base, shim.decoder = PM(), base, RefDecoder(d, d, C)
    shim.retrieve_weight, shim.label_weight, shim.finetune = 0.3, 0.3, True
    shim.noise_finetune, shim.query_graph_hop = False, 1
    shim.eval()
    with torch.no_grad():
        want = shim.forward(None, adj)
    mine = R.ToyGraphBase(None, C, d, 1, device="cpu", variant="graph")
    assert mine.retrieve_num == base.retrieve_num
    mine.add_entries(keys, values, labels)
    shim.toy_graph_base = mine
    with torch.no_grad():
        got = shim.forward(None, adj)
    assert got.shape == want.shape == (1, C) and float((got - want).abs().max()) < 1e-5


def test_node_fewshot_forward_with_our_store(reference, cpu_ops):  # noqa: F811
    """RAGraph_node_fewshot/RAGraph.py:47-83 unchanged (encode -> retrieve(emb, adj, noise) -> ... -> decode); our store
    derives the position-aware codes from the adjacency with the same seeded CPU randint as the reference's retrieve."""
    reference("RAGraph_node_fewshot")
    ref_tgb = sys.modules.get("ragraph_utils.ToyGraphBase") or importlib.import_module("ragraph_utils.ToyGraphBase")
    import layers.gcn as ref_gcn
    RefRAG = importlib.import_module("RAGraph").RAGraph
    g = torch.Generator().manual_seed(10)
    n, N, d, C = 22, 300, 32, 3
    adj = _sym_norm_adj(n, 0.15, g)
    emb = torch.randn(n, d, generator=g)
    keys = torch.nn.functional.normalize(torch.randn(N, d, generator=g), dim=-1)
    values = torch.randn(N, d, generator=g)
    labels = torch.nn.functional.one_hot(torch.randint(0, C, (N,), generator=g), C).float()
    positions = torch.rand(N, 10, generator=g)
    logits_table = torch.randn(C, C, generator=g)
    torch.manual_seed(11)
    dec = ref_gcn.GCN(d, C, "prelu")

    class PM:
        def encode(self, features, a):
            return emb

        def decode(self, hidden, a):
            with torch.no_grad():
                return dec((hidden, a))

    base = object.__new__(ref_tgb.ToyGraphBase)
    base.retrieve_num, base.noise_retrieve_num, base.num_anchors, base.dis_q = 5, 1, 10, 10
    base.structure_weight, base.semantic_weight = 0.001, 0.999
    base.resource_keys, base.resource_values, base.resource_labels, base.resource_positions = keys, values, labels, positions
    shim = object.__new__(RefRAG)
    torch.nn.Module.__init__(shim)
    shim.pretrain_model, shim.toy_graph_base = PM(), base
    shim.retrieve_weight, shim.label_weight, shim.finetune = 0.5, 0.5, True
    shim.noise_finetune, shim.query_graph_hop = False, 3
    shim.eval()
    with torch.no_grad():
        torch.manual_seed(12)
        want = shim.forward(None, adj, logits_table)
    mine = R.ToyGraphBase(None, C, d, 3, device="cpu", variant="node_fewshot")
    mine.retrieve_num = 5
    mine.add_entries(keys, values, labels, positions)
    shim.toy_graph_base = mine
    with torch.no_grad():
        torch.manual_seed(12)
        got = shim.forward(None, adj, logits_table)
    assert float((got - want).abs().max()) < 1e-5
    # and our own few-shot module with the reference's backbone object
    ours = R.RAGraphFewShot(PM(), mine, d, True, False, 3, 0.5, 0.5, graph_level=False).eval()
    with torch.no_grad():
        torch.manual_seed(12)
        got2 = ours(None, adj, logits_table)
    assert float((got2 - want).abs().max()) < 1e-5


def test_edge_forward_with_integration_patches(reference, cpu_ops):  # noqa: F811
    """modules/RAGraph.py:265-333 unchanged; scatter_sum, scatter_softmax and SimilarityFunctions swapped inside the module."""
    reference("RAGraph_edge", argv=["x", "--device", "cpu", "--data_path", "dataset/amazon"])
    from modules.RAGraph import RAGraph as RefEdge
    import modules.RAGraph as ref_mod
    from utils.parse_args import args
    g = torch.Generator().manual_seed(13)
    nu, ni, d, E = 40, 30, 16, 400
    n = nu + ni
    u = torch.randint(0, nu, (E,), generator=g); i = torch.randint(0, ni, (E,), generator=g) + nu
    edges = torch.cat([torch.stack([u, i], 1), torch.stack([i, u], 1)], 0)
    w = torch.rand(edges.shape[0], generator=g)
    times = torch.randint(0, 500, (edges.shape[0],), generator=g)
    X = torch.randn(n, d, generator=g)
    keys, values = torch.randn(n, d, generator=g), torch.randn(n, d, generator=g)
    fw = types.SimpleNamespace(num_users=nu, num_items=ni, phase="vanilla", use_RAG=True, use_noise=False, use_LoRA=False,
                               training=False, user_embedding=X[:nu], item_embedding=X[nu:], emb_gate=lambda x: x,
                               batch_size=32, retrieve_num=10, noise_retrieve_num=1, retrieve_weight=0.3,
                               resource_keys=keys, resource_values=values)
    fw._agg = types.MethodType(RefEdge._agg, fw)
    fw._relative_edge_time_encoding = types.MethodType(RefEdge._relative_edge_time_encoding, fw)
    with torch.no_grad():
        wu, wi = RefEdge.forward(fw, edges, w, times)
    want = torch.cat([wu, wi], 0)
    ref_mod.scatter_sum = R.scatter_sum
    ref_mod.scatter_softmax = lambda src, index, dim_size=None: R.scatter_softmax(src, index, dim_size=dim_size)
    ref_mod.SimilarityFunctions = R.SimilarityFunctions
    with torch.no_grad():
        gu, gi = RefEdge.forward(fw, edges, w, times)
    assert float((torch.cat([gu, gi], 0) - want).abs().max()) < 5e-6
    # the one-call form of the same forward
    got = R.edge_rag_forward(X, edges, w, keys, values, args.num_layers, 10, 32, 0.3, edge_times=times)
    assert float((got - want).abs().max()) < 5e-6


@pytest.mark.parametrize("variant", ["RAGraph_node", "RAGraph_node_fewshot", "RAGraph_graph"])
def test_library_build_matches_reference_build(reference, cpu_ops, variant):  # noqa: F811
    """The reference's own _build_toy_graph_base (augmentation x3 + inverse sampling of 10 nodes in the node variants;
    plain in the graph variant) and ours, from the same seed, produce the same library."""
    reference(variant)
    ref_tgb = sys.modules.get("ragraph_utils.ToyGraphBase") or importlib.import_module("ragraph_utils.ToyGraphBase")
    import layers.gcn as ref_gcn
    g = torch.Generator().manual_seed(2025)
    F_in, d, C = 12, 16, 3
    torch.manual_seed(3)
    layer = ref_gcn.GCN(F_in, d, "prelu")

    class PM:
        def inference(self, features, a):
            with torch.no_grad():
                return layer((features, a))

        encode = inference                                    # the few-shot variant calls the first layer `encode`

    graphs = []
    for n in (14, 9, 21):
        adj = _sym_norm_adj(n, 0.25, g)
        feats = torch.rand(n, F_in, generator=g)
        if variant == "RAGraph_graph":
            labels = torch.randint(0, C, (1,), generator=g)
        else:
            labels = torch.nn.functional.one_hot(torch.randint(0, C, (n,), generator=g), C).float()
        graphs.append((feats, adj, labels))

    kind = {"RAGraph_node": "node", "RAGraph_node_fewshot": "node_fewshot", "RAGraph_graph": "graph"}[variant]
    if kind == "node_fewshot":
        ref = ref_tgb.ToyGraphBase(PM(), C, d, 3, 5)
    else:
        ref = ref_tgb.ToyGraphBase(PM(), C, d, 3 if kind == "node" else 1)
    torch.manual_seed(99)
    for feats, adj, labels in graphs:
        ref._build_toy_graph_base(feats, adj, labels)

    mine = R.ToyGraphBase(PM(), C, d, 3 if kind != "graph" else 1, device="cpu", variant=kind, capacity=8,
                          label_dtype=torch.int64 if kind == "graph" else torch.float32)
    assert (mine.num_inverse_sample, mine.num_augment_scale) == (ref.num_inverse_sample, ref.num_augment_scale)
    torch.manual_seed(99)
    mine.build_toy_graph(graphs)

    assert len(mine) == ref.resource_keys.shape[0]
    if kind != "graph":
        assert len(mine) == 3 * (1 + 3) * 10                  # 3 graphs x (1 + 3 augmentations) x 10 sampled nodes
    finite = torch.isfinite(ref.resource_values).all(dim=1)    # augmented 0/1 adjacencies can have empty rows: 0/0 = NaN
    assert torch.equal(torch.isfinite(mine.resource_values).all(dim=1), finite)
    assert float((mine.resource_keys - ref.resource_keys).abs().max()) < 1e-6
    assert float((mine.resource_values[finite] - ref.resource_values[finite]).abs().max()) < 1e-5
    assert torch.equal(mine.resource_labels.to(ref.resource_labels.dtype), ref.resource_labels)
    if kind == "node_fewshot":
        assert float((mine.resource_positions - ref.resource_positions).abs().max()) < 1e-6


@pytest.mark.parametrize("scale,n_sample", [(0, 0), (0, 7), (1, 7)])
def test_edge_resource_graph_matches_reference_build(reference, cpu_ops, scale, n_sample):  # noqa: F811
    """modules/RAGraph.py:185-226 unchanged (finetune-phase and vanilla-phase knob settings) vs make_resource_graph."""
    reference("RAGraph_edge", argv=["x", "--device", "cpu", "--data_path", "dataset/amazon"])
    from modules.RAGraph import RAGraph as RefEdge
    from utils.parse_args import args
    g = torch.Generator().manual_seed(404)
    nu, ni, d, E = 35, 25, 16, 300
    n = nu + ni
    u = torch.randint(0, nu, (E,), generator=g); i = torch.randint(0, ni, (E,), generator=g) + nu
    adj_sp = torch.sparse_coo_tensor(torch.cat([torch.stack([u, i]), torch.stack([i, u])], 1),
                                     torch.rand(2 * E, generator=g), (n, n)).coalesce()
    X = torch.randn(n, d, generator=g)
    shim = types.SimpleNamespace(num_users=nu, num_items=ni, adj=adj_sp, edges=adj_sp._indices().t(),
                                 edge_norm=adj_sp._values(), resource_graph_radius=args.num_layers,
                                 num_augment_scale=scale, num_inverse_sample=n_sample, resource_keys=None, resource_values=None)
    shim._agg = types.MethodType(RefEdge._agg, shim)
    pre = types.SimpleNamespace(generate=lambda: (X[:nu], X[nu:]))
    with torch.no_grad():
        torch.manual_seed(55)
        RefEdge._make_resource_graph(shim, pre)
        torch.manual_seed(55)
        keys, values = R.make_resource_graph(X, shim.edges, shim.edge_norm, args.num_layers, num_augment_scale=scale,
                                             num_inverse_sample=n_sample, adj=adj_sp)
    assert keys.shape == shim.resource_keys.shape and values.shape == shim.resource_values.shape
    assert float((keys - shim.resource_keys).abs().max()) < 5e-6
    assert float((values - shim.resource_values).abs().max()) < 5e-6
    if scale or n_sample:                                     # adj omitted: the importance scores come from the transposed CSR
        torch.manual_seed(55)
        k2, v2 = R.make_resource_graph(X, shim.edges, shim.edge_norm, args.num_layers, num_augment_scale=scale,
                                       num_inverse_sample=n_sample)
        assert float((k2 - shim.resource_keys).abs().max()) < 5e-6 and float((v2 - shim.resource_values).abs().max()) < 5e-6


def test_process_tu_dataset_signature_and_result(reference, cpu_ops):  # noqa: F811
    """The reference's process_tu_dataset on a duck-typed Batch vs ours with the same call: features / labels identical,
    the CSR adjacency densifies to the reference's dense block-diagonal matrix."""
    reference("RAGraph_node")
    util = importlib.import_module("ragraph_utils.utility")
    g = torch.Generator().manual_seed(66)

    class Graph:
        def __init__(self, n, E):
            self.x = torch.rand(n, 9, generator=g)
            self.edge_index = torch.randint(0, n, (2, E), generator=g)

    class Batch(list):
        @property
        def num_graphs(self):
            return len(self)

        @property
        def num_features(self):
            return self[0].x.shape[1]

    data = Batch([Graph(6, 10), Graph(1, 0), Graph(11, 30)])
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                       # np.row_stack deprecation inside the reference
        rf, radj, rl = util.process_tu_dataset(data, 6)
    feats, csr, labs = R.process_tu_dataset(data, 6, device=torch.device("cpu"))
    assert torch.equal(feats, rf) and torch.equal(labs, rl)
    dense = torch.zeros(csr.n_rows, csr.n_cols)
    dense[csr.row_ids(), csr.col.long()] = csr.val
    assert float((dense - radj).abs().max()) <= 1e-7 and torch.equal(dense != 0, radj != 0)
    # and the CSR is accepted where the reference passes the dense matrix
    x = torch.randn(csr.n_rows, 8, generator=g)
    a = R.Propagation.aggregate_k_hop_features(csr, x, 2)
    b = O.aggregate_k_hop_features(radj, x, 2)
    assert float((a - b).abs().max()) < 1e-5


def test_fewshot_helpers_match_reference(reference, cpu_ops):  # noqa: F811
    """RAGraph_node_fewshot/ragraph_utils/utility.py:75-162 vs ragraph_b200.fewshot, same calls."""
    reference("RAGraph_node_fewshot")
    util = importlib.import_module("ragraph_utils.utility")
    F = R.fewshot
    g = torch.Generator().manual_seed(123)
    C, L_, n_support, n = 3, 12, 15, 40
    support_logits = torch.randn(n_support, L_, generator=g)
    support_labels = torch.arange(n_support) % C
    support_labels = support_labels[torch.randperm(n_support, generator=g)]
    logits = torch.randn(n, L_, generator=g)
    logits[3] = 0.0
    gold = torch.randint(0, C, (n,), generator=g)
    rm, ru = util.fewshot_mean(support_logits, support_labels)
    om, ou = F.fewshot_mean(support_logits, support_labels)
    assert torch.equal(ou, ru) and float((om - rm).abs().max()) < 1e-6
    assert float((F.fewshot_mean_logits(support_logits, support_labels) - util.fewshot_mean_logits(support_logits, support_labels)).abs().max()) < 1e-6
    rmap = util.fewshot_logits_map(support_logits, support_labels); omap = F.fewshot_logits_map(support_logits, support_labels)
    assert sorted(rmap) == sorted(omap) and all(float((omap[k] - rmap[k]).abs().max()) < 1e-6 for k in rmap)
    sim_r = util.fewshot_predict_logits(rm, logits); sim_o = F.fewshot_predict_logits(om, logits)
    assert sim_o.shape == sim_r.shape and float((sim_o - sim_r).abs().max()) < 2e-6
    assert torch.equal(F.fewshot_predict_labels_by_mean(om, logits)[4:], util.fewshot_predict_labels_by_mean(rm, logits)[4:])
    assert torch.equal(F.fewshot_predict_labels(support_logits, support_labels, logits)[4:],
                       util.fewshot_predict_labels(support_logits, support_labels, logits)[4:])
    lo = F.fewshot_predict_loss(support_logits, support_labels, logits, gold)
    lr = util.fewshot_predict_loss(support_logits, support_labels, logits, gold)
    assert abs(float(lo) - float(lr)) < 1e-6
    # non-contiguous label ids (e.g. {1, 4}): means follow the sorted unique labels; the 0..C-1 table refuses them
    odd = torch.tensor([4, 1, 4, 1, 1])
    om2, ou2 = F.fewshot_mean(support_logits[:5], odd); rm2, ru2 = util.fewshot_mean(support_logits[:5], odd)
    assert torch.equal(ou2, ru2) and float((om2 - rm2).abs().max()) < 1e-6
    with pytest.raises(KeyError):
        F.fewshot_mean_logits(support_logits[:5], odd)
    with pytest.raises(KeyError):
        util.fewshot_mean_logits(support_logits[:5], odd)
