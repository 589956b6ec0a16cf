"""Generate tests/golden/*.npz by running the UNMODIFIED reference hot-path functions.

TEST INFRASTRUCTURE (see oracle/ragraph_oracle.py header).  Run in the build container,
where /root/reference exists:

    python oracle/make_golden.py

The reference is pure Python/torch and cannot travel to the GPU box, so the vectors it
produces are committed as small fixtures together with this script.  Nothing in tests/,
smoke() or bench.py reads /root/reference at run time.

How the reference is made importable on CPU:
  * torch_geometric / torch_scatter are absent -> stub modules in sys.modules (only names
    the hot-path modules import at module scope; no stubbed function is on the path except
    scatter_softmax, which is third-party torch_scatter and restated in plain torch here);
  * ``.cuda()`` is hard-coded (ToyGraphBase.py:35-38, layers/gcn.py:28-29) ->
    ``torch.Tensor.cuda`` is patched to the identity for the duration of this script;
  * the five variants reuse the same top-level package names -> each variant is imported
    after purging those names from sys.modules.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
_VARIANT_PKGS = ("ragraph_utils", "layers", "models", "utils", "modules", "RAGraph",
                 "preprompt", "downprompt", "aug", "utility")


def _install_stubs():
    tg = types.ModuleType("torch_geometric")
    tgl = types.ModuleType("torch_geometric.loader"); tgl.DataLoader = object
    tgd = types.ModuleType("torch_geometric.datasets"); tgd.TUDataset = object
    tg.loader, tg.datasets = tgl, tgd
    sys.modules.update({"torch_geometric": tg, "torch_geometric.loader": tgl,
                        "torch_geometric.datasets": tgd})

    ts = types.ModuleType("torch_scatter")

    def scatter_softmax(src, index, dim_size=None):
        # torch_scatter.scatter_softmax restated (third-party, K10, out of hot-path scope)
        n = int(dim_size) if dim_size is not None else int(index.max()) + 1
        mx = torch.full((n,), -float("inf"), dtype=src.dtype).scatter_reduce(
            0, index, src, reduce="amax", include_self=True)
        e = torch.exp(src - mx[index])
        den = torch.zeros(n, dtype=src.dtype).scatter_add_(0, index, e)
        return e / den[index]

    ts.scatter_softmax = scatter_softmax
    sys.modules["torch_scatter"] = ts
    torch.Tensor.cuda = lambda self, *a, **k: self          # CPU-only container
    torch.nn.Module.cuda = lambda self, *a, **k: self


def _enter_variant(name, argv=None):
    for m in list(sys.modules):
        if m.split(".")[0] in _VARIANT_PKGS:
            del sys.modules[m]
    sys.path[:] = [p for p in sys.path if not p.startswith(REF)]
    sys.path.insert(0, os.path.join(REF, name))
    if argv is not None:
        sys.argv = argv


def _save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                                 for k, v in arrays.items()})
    print("wrote", os.path.relpath(path), {k: tuple(np.asarray(v).shape) for k, v in arrays.items()})


def _sym_norm_adj(n, p, gen):
    a = (torch.rand(n, n, generator=gen) < p).float()
    a = torch.triu(a, 1); a = a + a.t() + torch.eye(n)
    dinv = a.sum(1).pow(-0.5)
    return dinv[:, None] * a * dinv[None, :]


def gen_node():
    _enter_variant("RAGraph_node")
    from ragraph_utils.ToyGraphBase import ToyGraphBase
    from ragraph_utils.Propagation import Propagation
    from ragraph_utils.SimilarityFunctions import SimilarityFunctions
    from ragraph_utils.TaskDecoder import TaskDecoder
    from layers.gcn import GCN

    g = torch.Generator().manual_seed(20241017)
    Q, N, d, C = 37, 600, 32, 3
    q = torch.randn(Q, d, generator=g)
    q[5] = 0.0                                            # zero row -> eps clamp
    keys = torch.nn.functional.normalize(torch.randn(N, d, generator=g), dim=-1)
    keys[17] = keys[3]                                    # exact duplicate -> tie
    keys[44] = 0.0                                        # zero key
    values = torch.randn(N, d, generator=g)
    labels = torch.nn.functional.one_hot(torch.randint(0, C, (N,), generator=g), C).float()

    base = ToyGraphBase(None, C, d, 3)                    # real constructor (cuda() patched)
    base.resource_keys, base.resource_values, base.resource_labels = keys, values, labels
    S = SimilarityFunctions.calculate_cosine_similarity(q, keys)
    emb, lab = base.retrieve(q, None, False)
    torch.manual_seed(77)
    emb_n, lab_n = base.retrieve(q, None, True)
    torch.manual_seed(77)
    noise_idx = torch.randint(0, N, (Q, base.noise_retrieve_num))
    _save("node_retrieve", q=q, keys=keys, values=values, labels=labels, cosine=S,
          retrieve_num=base.retrieve_num, rag_embeddings=emb, rag_labels=lab,
          noise_indices=noise_idx, rag_embeddings_noise=emb_n, rag_labels_noise=lab_n)

    # Propagation
    n, Fdim = 50, 16
    adj = _sym_norm_adj(n, 0.08, g)
    x = torch.randn(n, Fdim, generator=g)
    outs = {f"out_k{k}": Propagation.aggregate_k_hop_features(adj, x, k) for k in (0, 1, 2, 3)}
    _save("propagation", adj=adj, x=x, **outs)

    # GCN layer (dense branch, unmodified forward)
    torch.manual_seed(5)
    layer = GCN(24, 16, 'prelu')
    with torch.no_grad():
        layer.bias.copy_(torch.randn(16, generator=g) * 0.1)
        layer.act.weight.fill_(0.25)
    seq = torch.randn(n, 24, generator=g)
    with torch.no_grad():
        out = layer((seq, adj.unsqueeze(0)))
    _save("gcn_layer", seq=seq, adj=adj, weight=layer.fc.weight.detach(), bias=layer.bias.detach(),
          alpha=layer.act.weight.detach(), out=out)

    # fusion arithmetic of RAGraph.forward (RAGraph_node/RAGraph.py:39-63), real method on a shim
    import importlib
    sys.modules.setdefault("utils", types.ModuleType("utils")).process = None
    RAG = importlib.import_module("RAGraph").RAGraph
    torch.manual_seed(11)
    dec = TaskDecoder(d, d, C)
    nq = 40
    adj_q = _sym_norm_adj(nq, 0.1, g)
    emb_q = torch.randn(nq, d, generator=g)

    class _PM:
        def inference(self, features, adj):
            return emb_q

    shim = object.__new__(RAG)
    torch.nn.Module.__init__(shim)
    shim.pretrain_model, shim.toy_graph_base, shim.decoder = _PM(), base, dec
    shim.retrieve_weight, shim.label_weight, shim.finetune = 0.5, 0.5, True
    shim.noise_finetune, shim.query_graph_hop = False, 3
    shim.eval()
    with torch.no_grad():
        logits = shim.forward(None, adj_q)
        shim.finetune = False
        vanilla = shim.forward(None, adj_q)
    _save("node_forward", emb_q=emb_q, adj_q=adj_q, keys=keys, values=values, labels=labels,
          w1=dec.fc1.weight.detach(), b1=dec.fc1.bias.detach(), w2=dec.fc2.weight.detach(),
          b2=dec.fc2.bias.detach(), logits=logits, vanilla=vanilla, retrieve_num=base.retrieve_num)


def gen_graph():
    _enter_variant("RAGraph_graph")
    from ragraph_utils.ToyGraphBase import ToyGraphBase
    g = torch.Generator().manual_seed(424242)
    N, d, C = 300, 32, 6
    keys = torch.randn(N, d, generator=g) * 0.3           # graph keys are means: NOT unit norm
    values = torch.randn(N, d, generator=g)
    labels = torch.nn.functional.one_hot(torch.randint(0, C, (N,), generator=g), C)   # int64
    q1 = torch.randn(d, generator=g)
    base = ToyGraphBase(None, C, d, 1)
    base.resource_keys, base.resource_values = keys, values
    base.resource_labels = torch.cat((torch.empty(0, C), labels), dim=0)   # promotion as in :126
    emb, lab = base.retrieve(q1, None, False)
    _save("graph_retrieve", q=q1, keys=keys, values=values, labels=base.resource_labels,
          retrieve_num=base.retrieve_num, rag_embeddings=emb, rag_labels=lab)


def gen_node_fewshot():
    _enter_variant("RAGraph_node_fewshot")
    from ragraph_utils.ToyGraphBase import ToyGraphBase
    from ragraph_utils.PositionAwareEncoder import PositionAwareEncoder
    g = torch.Generator().manual_seed(99)
    nq, N, d, C = 30, 400, 32, 3
    adj = _sym_norm_adj(nq, 0.12, g)
    q = torch.randn(nq, d, generator=g)
    keys = torch.nn.functional.normalize(torch.randn(N, d, generator=g), dim=-1)
    values = torch.randn(N, d, generator=g)
    labels = torch.nn.functional.one_hot(torch.randint(0, C, (N,), generator=g), C).float()
    positions = torch.rand(N, 10, generator=g)
    base = object.__new__(ToyGraphBase)
    base.retrieve_num, base.noise_retrieve_num = 5, 1
    base.num_anchors, base.dis_q = 10, 10
    base.structure_weight, base.semantic_weight = 0.001, 0.999
    base.resource_keys, base.resource_values = keys, values
    base.resource_labels, base.resource_positions = labels, positions
    torch.manual_seed(123)
    emb, lab = base.retrieve(q, adj, False)
    torch.manual_seed(123)
    search_positions = PositionAwareEncoder.encode_position_aware_code(adj, 10, 10)
    _save("fewshot_retrieve", q=q, search_positions=search_positions, keys=keys, positions=positions,
          values=values, labels=labels, retrieve_num=5, rag_embeddings=emb, rag_labels=lab)


def gen_edge():
    _enter_variant("RAGraph_edge", argv=["x", "--device", "cpu", "--data_path", "dataset/amazon"])
    from modules.RAGraph import RAGraph
    from modules.utils import scatter_sum
    from utils.parse_args import args
    g = torch.Generator().manual_seed(31337)
    nu, ni, d, E = 70, 50, 16, 900
    n = nu + ni
    u = torch.randint(0, nu, (E,), generator=g); i = torch.randint(0, ni, (E,), generator=g) + nu
    edges = torch.cat([torch.stack([u, i], 1), torch.stack([i, u], 1)], 0)           # symmetrised
    w = torch.rand(edges.shape[0], generator=g)
    X = torch.randn(n, d, generator=g)
    shim = types.SimpleNamespace(num_users=nu, num_items=ni)
    Y = RAGraph._agg(shim, X, edges, w)
    _save("edge_agg", X=X, edges=edges, w=w, num_nodes=n, Y=Y,
          scatter=scatter_sum(X[edges[:, 0]], edges[:, 1], dim=0, dim_size=n))

    # full forward (modules/RAGraph.py:265-333) on a shim, vanilla phase, no noise / LoRA
    keys = torch.randn(n, d, generator=g)
    values = torch.randn(n, d, generator=g)
    times = torch.randint(0, 1000, (edges.shape[0],), generator=g)
    fw = types.SimpleNamespace(
        num_users=nu, num_items=ni, phase="vanilla", use_RAG=True, use_noise=False, use_LoRA=False,
        training=False, user_embedding=X[:nu], item_embedding=X[nu:], emb_gate=lambda x: x,
        batch_size=32, retrieve_num=10, noise_retrieve_num=1, retrieve_weight=0.3,
        resource_keys=keys, resource_values=values)
    fw._agg = types.MethodType(RAGraph._agg, fw)
    fw._relative_edge_time_encoding = types.MethodType(RAGraph._relative_edge_time_encoding, fw)
    with torch.no_grad():
        time_norm = fw._relative_edge_time_encoding(edges, times)
        ur, ir = RAGraph.forward(fw, edges, w, times)
    # noisy training branch (:296,310,316-319): top-(k+1) + one torch.randint row per query and batch
    fw.use_noise, fw.training = True, True
    with torch.no_grad():
        torch.manual_seed(2024)
        urn, irn = RAGraph.forward(fw, edges, w, times)
    torch.manual_seed(2024)
    noise_idx = torch.cat([torch.randint(0, n, (min(s0 + 32, n) - s0, 1)) for s0 in range(0, n, 32)], 0)
    _save("edge_forward", X=X, edges=edges, w=w, times=times, time_norm=time_norm, keys=keys, values=values,
          num_layers=args.num_layers, batch_size=32, retrieve_num=10, retrieve_weight=0.3,
          out=torch.cat([ur, ir], 0), noise_seed=2024, noise_indices=noise_idx, out_noise=torch.cat([urn, irn], 0))


def gen_edge_eval():
    """Evaluation ranking (utils/metrics.py:96-118, 210-214): rating = U @ I^T, history items set to -1e8, torch.topk."""
    from utils.metrics import Metric
    from modules.LightGCN import LightGCN
    g = torch.Generator().manual_seed(2718)
    B, ni, d, k = 40, 300, 16, 20
    U = torch.randn(B, d, generator=g); I = torch.randn(ni, d, generator=g)
    hist = {u: torch.randperm(ni, generator=g)[: int(torch.randint(0, 60, (1,), generator=g))].tolist() for u in range(B)}
    m = object.__new__(Metric)
    with torch.no_grad():
        pred = LightGCN.rating(None, U, I).cpu()
    pred = m._mask_history_pos(pred, list(range(B)), hist)
    top_s, top_i = torch.topk(pred, k=k)
    rowptr = torch.zeros(B + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.tensor([len(hist[u]) for u in range(B)]), 0)
    cols = torch.tensor([x for u in range(B) for x in hist[u]], dtype=torch.int64)
    _save("edge_eval", U=U, I=I, hist_rowptr=rowptr, hist_items=cols, k=k, top_scores=top_s, top_items=top_i)


def _fewshot_forward(variant, graph_level, seed, rw, lw, hop):
    """Unmodified RAGraph.forward of a few-shot variant on a shim (RAGraph_node_fewshot/RAGraph.py:47-83,
    RAGraph_graph_fewshot/RAGraph.py:46-91); encode returns fixed embeddings, decode is the variant's own GCN layer."""
    _enter_variant(variant)
    if graph_level:
        # RAGraph_graph_fewshot/ragraph_utils/__init__.py:7 imports a module that is not in the repo, and
        # FewShotBase (:2) is not on the forward path: stub the missing names, nothing of them is called
        fu = types.ModuleType("ragraph_utils.fewshot_utility")
        for nm in ("fewshot_predict_labels_by_mean", "fewshot_mean_logits", "fewshot_predict_logits", "fewshot_predict_labels"):
            setattr(fu, nm, None)
        sys.modules["ragraph_utils.fewshot_utility"] = fu
    import importlib
    from ragraph_utils.ToyGraphBase import ToyGraphBase
    from layers.gcn import GCN
    RAG = importlib.import_module("RAGraph").RAGraph
    g = torch.Generator().manual_seed(seed)
    nq, N, d, C = 26, 350, 32, (3 if not graph_level else 2)
    adj = _sym_norm_adj(nq, 0.12, g)
    emb_q = torch.randn(nq, d, generator=g)
    keys = torch.nn.functional.normalize(torch.randn(N, d, generator=g), dim=-1)
    values = torch.randn(N, d, generator=g)
    labels = torch.nn.functional.one_hot(torch.randint(0, C, (N,), generator=g), C).float()
    logits_table = torch.randn(C, C, generator=g)
    torch.manual_seed(seed)
    dec = GCN(d, C, 'prelu')
    with torch.no_grad():
        dec.bias.copy_(torch.randn(C, generator=g) * 0.1)

    class _PM:
        def encode(self, features, a):
            return emb_q

        def decode(self, hidden, a):
            return dec((hidden, a))

    base = object.__new__(ToyGraphBase)
    base.retrieve_num, base.noise_retrieve_num = (5, 1) if not graph_level else (min(3, C + 1), 1)
    base.resource_keys, base.resource_values, base.resource_labels = keys, values, labels
    extra = {}
    if not graph_level:
        from ragraph_utils.PositionAwareEncoder import PositionAwareEncoder
        base.num_anchors, base.dis_q = 10, 10
        base.structure_weight, base.semantic_weight = 0.001, 0.999
        base.resource_positions = torch.rand(N, 10, generator=g)
        extra["positions"] = base.resource_positions
    shim = object.__new__(RAG)
    torch.nn.Module.__init__(shim)
    shim.pretrain_model, shim.toy_graph_base = _PM(), base
    shim.retrieve_weight, shim.label_weight, shim.finetune = rw, lw, True
    shim.noise_finetune, shim.query_graph_hop = False, hop
    shim.eval()
    with torch.no_grad():
        torch.manual_seed(seed + 1)                      # PositionAwareEncoder draws random anchors
        out = shim.forward(None, adj, logits_table)
        shim.finetune = False
        torch.manual_seed(seed + 1)
        vanilla = shim.forward(None, adj, logits_table)
        if not graph_level:
            torch.manual_seed(seed + 1)
            extra["search_positions"] = PositionAwareEncoder.encode_position_aware_code(adj, 10, 10)
    _save("fewshot_forward_" + ("graph" if graph_level else "node"), emb_q=emb_q, adj=adj, keys=keys, values=values,
          labels=labels, mean_fewshot_logits=logits_table, dec_weight=dec.fc.weight.detach(), dec_bias=dec.bias.detach(),
          dec_alpha=dec.act.weight.detach(), retrieve_num=base.retrieve_num, retrieve_weight=rw, label_weight=lw,
          hop=hop, out=out, vanilla=vanilla, **extra)


def gen_fewshot_forward():
    _fewshot_forward("RAGraph_node_fewshot", False, 515, 0.5, 0.5, 3)
    _fewshot_forward("RAGraph_graph_fewshot", True, 616, 0.3, 0.8, 1)


def gen_downprompt():
    """Downstream prompt heads (section 8 row a9).  The reference allocates its class-slot scratch with
    torch.FloatTensor(...) -- UNINITIALISED memory -- and averages over it (RAGraph_node/downprompt.py:61,77;
    RAGraph_graph/downprompt.py:60,93,102): its output is defined only when the unwritten slots are zero.  For the
    duration of these calls torch.FloatTensor(sizes...) is patched to return zero-filled storage; nothing else of the
    reference is touched."""
    real_ft = torch.FloatTensor

    def zero_ft(*a, **k):
        if a and all(isinstance(x, int) for x in a):
            return torch.zeros(*a)
        return real_ft(*a, **k)

    g = torch.Generator().manual_seed(8128)
    try:
        torch.FloatTensor = zero_ft
        # ---- node variant: ELU(weight * emb), prototypes over n//2 slots, softmax of cosines
        _enter_variant("RAGraph_node")
        import importlib
        dpn = importlib.import_module("downprompt")
        n, d, C = 60, 24, 3
        feature = torch.randn(1, n, d, generator=g)
        labels = torch.randint(0, C, (n,), generator=g)
        labels[:5] = torch.tensor([0, 1, 2, 0, 1])
        seq = torch.randn(n, d, generator=g)
        seq[7] = 0.0                                      # zero row: cosine eps clamp
        p = torch.zeros(1, d)
        torch.manual_seed(3)
        m = dpn.downprompt(p, p, p, d, C, feature, labels)
        with torch.no_grad():
            ave0 = m.ave.clone()
            prompted = m.downprompt(seq)
            probs_eval = m.forward(seq, train=0)
            probs_train = m.forward(seq, train=1)
            ave1 = m.ave.clone()
        _save("downprompt_node", feature=feature, labels=labels, seq=seq, weight=m.downprompt.weight.detach(),
              ave_init=ave0, prompted=prompted, probs_eval=probs_eval, probs_train=probs_train, ave_train=ave1)

        # ---- graph variant: weight * emb, per-graph sum readout, prototypes over n slots, log_softmax of cosines
        _enter_variant("RAGraph_graph")
        dpg = importlib.import_module("downprompt")
        sizes = torch.tensor([5, 1, 9, 3, 12, 7, 2, 6])
        nn_, C = int(sizes.sum()), 6
        seq = torch.randn(nn_, d, generator=g)
        torch.manual_seed(4)
        mg = dpg.downprompt(p, p, p, d, C)
        glabels = torch.tensor([0, 5, 2, 2, 1, 3, 4, 0])
        with torch.no_grad():
            gemb = mg.forward(seq, sizes)
            ave = dpg.averageemb(glabels, gemb, C)
            logp = dpg.predict(sizes.shape[0], C, gemb, ave)
        _save("downprompt_graph", seq=seq, graph_sizes=sizes, weight=mg.downprompt.weight.detach(), graph_emb=gemb,
              graph_labels=glabels, ave=ave, log_probs=logp)
    finally:
        torch.FloatTensor = real_ft


def gen_library_build():
    """Library construction with the RNG-driven steps disabled (num_augment_scale = 0, num_inverse_sample = 0): the
    unmodified _build_toy_graph_base of the node and graph variants and the edge _make_resource_graph."""
    g = torch.Generator().manual_seed(60606)
    d, C = 24, 3
    graphs = []
    for n in (17, 9, 30):
        graphs.append((_sym_norm_adj(n, 0.2, g), torch.randn(n, d, generator=g),
                       torch.nn.functional.one_hot(torch.randint(0, C, (n,), generator=g), C).float(),
                       torch.randint(0, C, (1,), generator=g)))

    class _PM:
        def __init__(self):
            self.emb = None

        def inference(self, features, adj):
            return self.emb

    out = {}
    for variant in ("RAGraph_node", "RAGraph_graph"):
        _enter_variant(variant)
        from ragraph_utils.ToyGraphBase import ToyGraphBase
        pm = _PM()
        base = ToyGraphBase(pm, C, d, 3 if variant == "RAGraph_node" else 1)      # real constructor (cuda() patched)
        base.num_augment_scale, base.num_inverse_sample = 0, 0
        torch.manual_seed(1)                                                       # PositionAwareEncoder anchors (unused)
        for adj, emb, nl, gl in graphs:
            pm.emb = emb
            base._build_toy_graph_base(None, adj, nl if variant == "RAGraph_node" else gl)
        tag = "node" if variant == "RAGraph_node" else "graph"
        out[f"{tag}_keys"], out[f"{tag}_values"], out[f"{tag}_labels"] = base.resource_keys, base.resource_values, base.resource_labels
        out[f"{tag}_hop"] = base.toy_graph_hop
    for i, (adj, emb, nl, gl) in enumerate(graphs):
        out[f"adj{i}"], out[f"emb{i}"], out[f"node_labels{i}"], out[f"graph_label{i}"] = adj, emb, nl, gl

    # edge: _make_resource_graph on a shim
    _enter_variant("RAGraph_edge", argv=["x", "--device", "cpu", "--data_path", "dataset/amazon"])
    from modules.RAGraph import RAGraph
    from utils.parse_args import args
    nu, ni, de, E = 40, 30, 16, 400
    n = nu + ni
    u = torch.randint(0, nu, (E,), generator=g); i = torch.randint(0, ni, (E,), generator=g) + nu
    adj_sp = torch.sparse_coo_tensor(torch.cat([torch.stack([u, i]), torch.stack([i, u])], 1),
                                     torch.rand(2 * E, generator=g), (n, n)).coalesce()
    X = torch.randn(n, de, generator=g)
    shim = types.SimpleNamespace(num_users=nu, num_items=ni, adj=adj_sp, edges=adj_sp._indices().t(),
                                 edge_norm=adj_sp._values(), resource_graph_radius=args.num_layers,
                                 num_augment_scale=0, num_inverse_sample=0, resource_keys=None, resource_values=None)
    shim._agg = types.MethodType(RAGraph._agg, shim)
    pre = types.SimpleNamespace(generate=lambda: (X[:nu], X[nu:]))
    with torch.no_grad():
        RAGraph._make_resource_graph(shim, pre)
    out.update(edge_X=X, edge_edges=shim.edges, edge_w=shim.edge_norm, edge_radius=args.num_layers,
               edge_keys=shim.resource_keys, edge_values=shim.resource_values)
    _save("library_build", n_graphs=len(graphs), **out)


def gen_graph_forward():
    """Unmodified RAGraph_graph/RAGraph.py:48-75 forward on a shim (one query graph -> [1, C] class probabilities)."""
    _enter_variant("RAGraph_graph")
    import importlib
    from ragraph_utils.ToyGraphBase import ToyGraphBase
    from ragraph_utils.TaskDecoder import TaskDecoder
    RAG = importlib.import_module("RAGraph").RAGraph
    g = torch.Generator().manual_seed(777)
    nq, N, d, C = 23, 260, 32, 6
    adj = _sym_norm_adj(nq, 0.15, g)
    emb_q = torch.randn(nq, d, generator=g)
    keys = torch.randn(N, d, generator=g) * 0.3
    values = torch.randn(N, d, generator=g)
    labels = torch.nn.functional.one_hot(torch.randint(0, C, (N,), generator=g), C)
    base = ToyGraphBase(None, C, d, 1)
    base.resource_keys, base.resource_values = keys, values
    base.resource_labels = torch.cat((torch.empty(0, C), labels), dim=0)
    torch.manual_seed(12)
    dec = TaskDecoder(d, d, C)

    class _PM:
        def inference(self, features, a):
            return emb_q

    shim = object.__new__(RAG)
    torch.nn.Module.__init__(shim)
    shim.pretrain_model, shim.toy_graph_base, shim.decoder = _PM(), base, dec
    shim.retrieve_weight, shim.label_weight, shim.finetune = 0.3, 0.3, True
    shim.noise_finetune, shim.query_graph_hop = False, 1
    shim.eval()
    with torch.no_grad():
        logits = shim.forward(None, adj)
        shim.finetune = False
        vanilla = shim.forward(None, adj)
    _save("graph_forward", emb_q=emb_q, adj=adj, keys=keys, values=values, labels=base.resource_labels,
          w1=dec.fc1.weight.detach(), b1=dec.fc1.bias.detach(), w2=dec.fc2.weight.detach(), b2=dec.fc2.bias.detach(),
          logits=logits, vanilla=vanilla, retrieve_num=base.retrieve_num)


def gen_inverse_sampling():
    """InverseSampling.compute_sample_prob: dense variant (RAGraph_node/ragraph_utils/InverseSampling.py:6-56) on an
    adjacency with a dangling node, sparse variant (RAGraph_edge/modules/ragraph_utils/InverseSampling.py:6-69)."""
    g = torch.Generator().manual_seed(4242)
    _enter_variant("RAGraph_node")
    from ragraph_utils.InverseSampling import InverseSampling as ISdense
    n = 40
    a = (torch.rand(n, n, generator=g) < 0.1).float()
    a = torch.triu(a, 1); a = a + a.t()
    a[7, :] = 0.0                                           # node 7 has no out-links (but keeps in-links): dangling row
    with torch.no_grad():
        pr = ISdense.pagerank_algorithm(a.clone())
        dc = ISdense.degree_centrality_algorithm(a.clone())
        sp = ISdense.compute_sample_prob(a.clone())
    adj_norm = _sym_norm_adj(25, 0.2, g)
    with torch.no_grad():
        sp_norm = ISdense.compute_sample_prob(adj_norm.clone())
    _enter_variant("RAGraph_edge", argv=["x", "--device", "cpu", "--data_path", "dataset/amazon"])
    from modules.ragraph_utils.InverseSampling import InverseSampling as ISsparse
    nu, ni, E = 30, 20, 200
    m = nu + ni
    u = torch.randint(0, nu, (E,), generator=g); i = torch.randint(0, ni - 1, (E,), generator=g) + nu   # last item isolated
    adj_sp = torch.sparse_coo_tensor(torch.cat([torch.stack([u, i]), torch.stack([i, u])], 1),
                                     torch.rand(2 * E, generator=g), (m, m)).coalesce()
    with torch.no_grad():
        sp_sparse = ISsparse.compute_sample_prob(adj_sp)
        pr_sparse = ISsparse.pagerank_algorithm(adj_sp)
    _save("inverse_sampling", adj=a, pagerank=pr, degree_centrality=dc, sample_prob=sp, adj_norm=adj_norm,
          sample_prob_norm=sp_norm, sparse_indices=adj_sp._indices(), sparse_values=adj_sp._values(), sparse_n=m,
          pagerank_sparse=pr_sparse, sample_prob_sparse=sp_sparse)


if __name__ == "__main__":
    assert os.path.isdir(REF), "run in the build container (needs /root/reference)"
    _install_stubs()
    gen_node(); gen_graph(); gen_node_fewshot(); gen_edge(); gen_edge_eval()
    gen_fewshot_forward(); gen_downprompt(); gen_library_build(); gen_graph_forward(); gen_inverse_sampling()
