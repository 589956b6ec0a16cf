"""CPU-side checks: the C-ABI library builds/loads and exports every declared symbol; host logic."""
import os
import re

import pytest
import torch

import ragraph_b200
from ragraph_b200 import _lib
from ragraph_b200.sharded import owner_of, shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ragraph_b200.h")).read()
    return sorted(set(re.findall(r"RAG_API\s+[\w\s\*]+?\b(rag_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound():
    if not os.path.exists(_lib.LIB_PATH):
        from ragraph_b200 import build
        build.build()
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ragraph_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.rag_abi_version() == _lib.ABI_VERSION == 2
    assert lib.rag_status_string(-3) == b"RAG_EUNSUPPORTED"


def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ragraph_b200.ops.cosine_topk(torch.randn(2, 8), torch.randn(9, 8), 2)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ragraph_b200.ops.gather_rows(torch.randn(4, 8), torch.zeros(2, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ragraph_b200.Propagation.aggregate_k_hop_features(torch.eye(3), torch.randn(3, 4), 1)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "ragraph_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


@pytest.mark.parametrize("n,world", [(10, 1), (10, 3), (7, 8), (100_000_000, 8), (5, 5), (0, 2)])
def test_shard_bounds_partition(n, world):
    spans = [shard_bounds(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and b >= a
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1
    if 0 < n <= 1000:
        idx = torch.arange(n)
        own = owner_of(idx, n, world)
        for r, (a, b) in enumerate(spans):
            assert torch.all(own[a:b] == r)


def test_fake_kernels_shape_inference():
    q = torch.empty(5, 16, device="meta"); keys = torch.empty(40, 16, device="meta")
    s, i = torch.ops.ragraph.cosine_topk(q, keys, 3)
    assert s.shape == (5, 3) and i.dtype == torch.int64
    g = torch.ops.ragraph.gather_rows(keys, torch.empty(5, 3, dtype=torch.int64, device="meta"))
    assert g.shape == (5, 3, 16)


def test_csr_transpose_and_row_normalized_host_logic():
    """CSRGraph.transpose / row_normalized are torch index plumbing (no kernel): check them against scipy on CPU."""
    import numpy as np
    import scipy.sparse as sp
    from ragraph_b200.csr import CSRGraph
    rng = np.random.default_rng(3)
    n, m, nnz = 50, 37, 400
    rows = np.sort(rng.integers(0, n, nnz)); cols = rng.integers(0, m, nnz); vals = rng.random(nnz).astype(np.float32) + 0.1
    rowptr = np.zeros(n + 1, np.int64); np.add.at(rowptr, rows + 1, 1); rowptr = np.cumsum(rowptr)
    g = CSRGraph(torch.from_numpy(rowptr), torch.from_numpy(cols.astype(np.int32)), torch.from_numpy(vals), n, m)
    A = sp.csr_matrix((vals, cols, rowptr), shape=(n, m))
    t = g.transpose()
    assert t.n_rows == m and t.n_cols == n and t.transpose() is g
    At = sp.csr_matrix((t.val.numpy(), t.col.numpy(), t.rowptr.numpy()), shape=(m, n))
    assert np.array_equal(At.toarray(), A.T.toarray())
    # entries of every transposed row are ordered by original row (fixed summation order in the backward SpMM)
    for r in range(m):
        seg = t.col[t.rowptr[r]:t.rowptr[r + 1]].numpy()
        assert np.all(np.diff(seg) >= 0)
    rn = g.row_normalized()
    dense = A.toarray(); deg = dense.sum(1, keepdims=True)
    with np.errstate(invalid="ignore", divide="ignore"):
        want = dense / deg
    got = sp.csr_matrix((rn.val.numpy(), rn.col.numpy(), rn.rowptr.numpy()), shape=(n, m)).toarray()
    ok = deg[:, 0] > 0
    assert np.allclose(got[ok], want[ok], rtol=1e-6)


def test_library_save_load_roundtrip_and_shards(tmp_path):
    """ToyGraphBase.save / load is host I/O plus tensor copies: bit-exact round trip, int64 labels,
    position codes of the few-shot variant, and loading one key-row shard with its global offset."""
    from ragraph_b200 import ToyGraphBase
    g = torch.Generator().manual_seed(4)
    n, d, C = 1000, 24, 5
    for variant, ldt in (("node", torch.float32), ("node_fewshot", torch.int64)):
        st = ToyGraphBase(None, C, d, 3, device="cpu", variant=variant, capacity=16, label_dtype=ldt)
        keys, vals = torch.randn(n, d, generator=g), torch.randn(n, d, generator=g)
        vals[3, 0] = -0.0
        labs = torch.nn.functional.one_hot(torch.randint(0, C, (n,), generator=g), C).to(ldt)
        pos = torch.rand(n, 10, generator=g) if variant == "node_fewshot" else None
        st.add_entries(keys[:600], vals[:600], labs[:600], None if pos is None else pos[:600])
        st.add_entries(keys[600:], vals[600:], labs[600:], None if pos is None else pos[600:])
        st.retrieve_num = 7
        p = str(tmp_path / variant)
        st.save(p, chunk_rows=333)
        back = ToyGraphBase.load(p, device="cpu", chunk_rows=100)
        assert len(back) == n and back.variant == variant and back.retrieve_num == 7 and back.shard_lo == 0
        assert back.resource_labels.dtype == ldt
        for a, b in ((back.resource_keys, keys), (back.resource_values, vals), (back.resource_labels, labs)):
            assert torch.equal(a.view(torch.uint8).reshape(-1) if a.dtype.is_floating_point else a.reshape(-1),
                               b.view(torch.uint8).reshape(-1) if b.dtype.is_floating_point else b.reshape(-1))
        if pos is not None:
            assert torch.equal(back.resource_positions, pos)
        lo, hi = shard_bounds(n, 3, 1)
        part = ToyGraphBase.load(p, device="cpu", rows=(lo, hi))
        assert len(part) == hi - lo and part.shard_lo == lo and torch.equal(part.resource_keys, keys[lo:hi])
    with pytest.raises(RuntimeError, match="outside"):
        ToyGraphBase.load(p, device="cpu", rows=(10, n + 1))


def test_process_graph_batch_matches_reference_dense_adjacency():
    """CSR-direct preprocessing vs the oracle restatement of process_tu_dataset + normalize_adj:
    asymmetric edges, duplicate edges, an isolated node, several graphs."""
    import numpy as np
    from oracle import ragraph_oracle as O
    from ragraph_b200 import process_graph_batch
    g = torch.Generator().manual_seed(12)
    xs, eis = [], []
    for n, E in ((7, 15), (1, 0), (12, 40), (5, 6)):
        xs.append(torch.rand(n, 9, generator=g))
        e = torch.randint(0, n, (2, E), generator=g)
        if E >= 6:
            e[:, 1] = e[:, 0]                                   # duplicate edge -> multiplicity 2
        eis.append(e)
    feats, csr, labs = process_graph_batch(xs, eis, 6)
    rf, radj, rl = O.process_tu_arrays([x.numpy() for x in xs], [e.numpy() for e in eis], 6)
    assert torch.equal(feats, rf) and torch.equal(labs, rl)
    dense = torch.zeros(csr.n_rows, csr.n_cols)
    rows = csr.row_ids()
    dense[rows, csr.col.long()] = csr.val
    assert torch.equal(dense != 0, radj != 0)
    assert float((dense - radj).abs().max()) <= 1e-7


def test_bench_stdout_guard_keeps_one_line():
    """bench.py's fd-level guard: C-level and Python-level writes to stdout before restore() land on stderr."""
    import subprocess
    import sys
    code = ("import os, sys; sys.path.insert(0, %r); import bench; g = bench.StdoutToStderr(); os.write(1, b'NCCL version x\\n'); "
            "print('python noise'); g.restore(); print('{\"ok\": 1}')" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout == '{"ok": 1}\n'
    assert "NCCL version x" in r.stderr and "python noise" in r.stderr


def test_retrieval_dispatch_rules():
    """rag_cosine_topk_plan (host arithmetic, 148 SMs assumed without a device): which kernel serves which shape.  Pins the
    measured dispatch rules of DESIGN 3.1-3.2b: two-pass mode on short key streams, the TS kernel (pre-pass + cross-split sweep on
    the SMs the grid leaves idle) from 1 024 key tiles per CTA, SS in between for d > 128, more key splits for k > 26."""
    L = _lib
    F16R, BF16R, BF16, FP32 = L.SIM_F16_REFINE, L.SIM_BF16_REFINE, L.SIM_BF16, L.SIM_FP32
    plan = L.cosine_topk_plan
    # the headline shapes: query-stationary kernel, 16 query tiles x 9 key splits = 144 workers + 4 sweeping CTAs, pre-pass = 1/64
    for N in (100_000_000, 50_000_000, 25_000_000, 12_500_000):
        p = plan(4096, N, 128, 10, BF16R)
        assert (p["kernel"], p["q_tiles"], p["key_splits"], p["sweep_ctas"], p["list_len"]) == ("ts", 16, 9, 4, 16), p
        assert p["prepass_tiles"] == min(p["tiles_per_cta"] // 64, 192)
    assert plan(4096, 10_000_000, 256, 10, F16R)["kernel"] == "ts"
    assert plan(4096, 100_000_000, 128, 10, BF16R, L.SIM_WIDE_LISTS)["list_len"] == 32
    # the reference's own library sizes: two passes (group maxima + collect) instead of list warm-up -- exact modes only
    for shape in ((2708, 10832, 256, 4), (4096, 240_000, 64, 10), (300, 20_000, 128, 10), (4096, 1_000_000, 128, 10),
                  (32768, 240_000, 64, 50), (4096, 240_000, 64, 21)):
        assert plan(*shape, F16R)["kernel"] == "two_pass", shape
        assert plan(*shape, BF16)["kernel"] in ("ss", "ts"), shape                    # raw modes keep their lists
    assert plan(4096, 240_000, 64, 10, F16R, 0, True)["kernel"] == "ss"               # exclusion lists: no two-pass, no sweep
    assert plan(4096, 100_000_000, 128, 10, F16R, 0, True)["sweep_ctas"] == 0
    # d > 128: the second pass costs twice the MMA time -> two-pass only up to 192 tiles per CTA, SS up to 1 024, then TS
    assert plan(4096, 200_000, 256, 10, F16R)["kernel"] == "two_pass"
    assert plan(4096, 500_000, 256, 10, F16R)["kernel"] == "ss"
    assert plan(4096, 2_000_000, 256, 10, F16R)["kernel"] == "ts"
    assert plan(4096, 2_000_000, 128, 10, F16R)["kernel"] == "ts"
    # k in (26, 128]: the lists of a row hold >= 4k entries together (more key splits, CTAs in waves)
    p = plan(32768, 240_000, 64, 50, BF16)
    assert p["list_len"] == 32 and p["key_splits"] * 32 >= 4 * 50 and p["q_tiles"] * p["key_splits"] > 148, p
    assert plan(4096, 2_000_000, 128, 100, BF16)["key_splits"] >= 13
    # outside the tensor-core range: the fp32 kernel
    assert plan(300, 70_000, 266, 5, F16R)["kernel"] == "fp32"
    assert plan(300, 70_000, 256, 20, F16R)["kernel"] == "fp32"
    assert plan(300, 70_000, 128, 10, FP32)["kernel"] == "fp32"
    # options are honoured and restorable
    try:
        L.tc_set_option("twopass", 0)
        assert plan(2708, 10832, 256, 4, F16R)["kernel"] == "ss"
        L.tc_set_option("gshare", 0)
        assert plan(4096, 12_500_000, 128, 10, BF16R)["sweep_ctas"] == 0
        L.tc_set_option("variant", 1)
        assert plan(4096, 12_500_000, 128, 10, BF16R)["kernel"] == "ss"
    finally:
        for name in ("twopass", "gshare", "variant"):
            L.tc_set_option(name, -1)
    assert plan(2708, 10832, 256, 4, F16R)["kernel"] == "two_pass"
