"""Secondary measurements on one B200 (not the driver's bench line): the other BASELINE configs and the
"kernel to beat on the same box" -- the reference's own torch calls executed by stock torch CUDA.

    python tools/bench_extra.py [section ...]      sections: cfg3 gather torch_gpu small spmm_variants
Writes one JSON object per line to stdout.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

import ragraph_b200 as R
from ragraph_b200 import _lib as L, ops

DEV = "cuda"
PEAKS = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return ms[len(ms) // 2]


def normal_rows(n, d, seed, normalize=True):
    out = torch.empty(n, d, device=DEV)
    g = torch.Generator(device=DEV).manual_seed(seed)
    for a in range(0, n, 4_000_000):
        b = min(n, a + 4_000_000)
        x = torch.randn(b - a, d, generator=g, device=DEV)
        out[a:b] = F.normalize(x, dim=-1) if normalize else x
    return out


def emit(**kw):
    print(json.dumps(kw), flush=True)


def cfg3():
    """10 M keys x d=256, 4 096-query batches, top-10, 1 GPU (BASELINE config 3)."""
    N, d, Q, k = 10_000_000, 256, 4096, 10
    keys = normal_rows(N, d, 1234); q = torch.randn(Q, d, device=DEV)
    inv = ops.row_inv_norm(keys); shadow = ops.rows_to_bf16(keys, True)
    for mode, name in ((L.SIM_BF16_REFINE, "bf16 filter + f32 refine (exact)"), (L.SIM_BF16, "bf16 raw")):
        ms = timeit(lambda: ops.cosine_topk(q, keys, k, key_inv_norm=inv, keys_bf16=shadow, mode=mode), 10)
        tf = 2.0 * Q * N * d / ms / 1e9
        emit(section="cfg3", workload=f"top-{k} cosine, {N} keys d={d}, Q={Q}", mode=name, ms=ms, qps=Q / ms * 1e3,
             tflops=tf, frac_of_bf16_peak=tf / PEAKS["bf16_tflops"])
    s3, i3 = ops.cosine_topk(q[:256], keys, k, key_inv_norm=inv, keys_bf16=shadow, mode=L.SIM_BF16_REFINE)
    s0, i0 = ops.cosine_topk(q[:256], keys, k, key_inv_norm=inv)
    emit(section="cfg3", check="mode3 == fp32 path on 256 queries", max_score_diff=float((s3 - s0).abs().max()),
         idx_equal_frac=float((i3 == i0).float().mean()))


def gather():
    """retrieved-subgraph gather, ogbn-products shape: Q=2.4 M, k=10, d=256 (BASELINE config 5)."""
    N, d, Q, k = 2_449_029, 256, 2_400_000, 10
    table = torch.randn(N, d, device=DEV)
    idx = torch.randint(0, N, (Q, k), device=DEV)
    ms = timeit(lambda: ops.gather_rows(table, idx), 5)
    alg = Q * k * (8 + 2 * d * 4)
    emit(section="gather", kernel="gather_rows", workload=f"Q={Q} k={k} d={d}", ms=ms, gbs=alg / ms / 1e6,
         frac_of_hbm=alg / ms / 1e6 / PEAKS["hbm_gbs"], algorithmic_bytes=alg)
    ms_t = timeit(lambda: table[idx], 5)
    emit(section="gather", kernel="torch index (stock)", ms=ms_t, gbs=alg / ms_t / 1e6)
    ms = timeit(lambda: ops.gather_reduce(table, idx, L.REDUCE_MEAN), 5)
    alg2 = Q * k * (8 + d * 4) + Q * d * 4
    emit(section="gather", kernel="gather_reduce(mean)", ms=ms, gbs=alg2 / ms / 1e6, frac_of_hbm=alg2 / ms / 1e6 / PEAKS["hbm_gbs"],
         algorithmic_bytes=alg2)
    ms_t = timeit(lambda: table[idx].mean(1), 5)
    emit(section="gather", kernel="torch index + mean (stock)", ms=ms_t, gbs=alg2 / ms_t / 1e6)
    out = ops.gather_rows(table, idx[:100000])
    emit(section="gather", check="bit exact vs torch index", ok=bool(torch.equal(out, table[idx[:100000]])))


def torch_gpu():
    """The reference's torch calls on the B200 (cuBLAS fp32 sgemm + at::topk + index; dense / scatter propagation)."""
    d, Q, k = 128, 4096, 10
    for N in (1_000_000, 2_000_000):
        keys = normal_rows(N, d, 1234); vals = torch.randn(N, d, device=DEV); q = torch.randn(Q, d, device=DEV)

        def ref():
            s = torch.matmul(F.normalize(q, p=2, dim=-1), F.normalize(keys, p=2, dim=-1).t())
            _, i = torch.topk(s, k, largest=True, sorted=True)
            return vals[i]
        ms = timeit(ref, 5, 2)
        emit(section="torch_gpu", what="reference retrieve (SimilarityFunctions + torch.topk + index) on stock torch CUDA",
             N=N, d=d, Q=Q, ms=ms, qps_at_N=Q / ms * 1e3, qps_extrapolated_100M=Q / (ms * 100_000_000 / N) * 1e3)
        inv = ops.row_inv_norm(keys); shadow = ops.rows_to_bf16(keys, True)
        ms_o = timeit(lambda: ops.gather_rows(vals, ops.cosine_topk(q, keys, k, key_inv_norm=inv, keys_bf16=shadow, mode=3)[1]), 5, 2)
        emit(section="torch_gpu", what="ours (mode 3) same shape", N=N, ms=ms_o, speedup=ms / ms_o)
        del keys, vals, inv, shadow
    # propagation: edge _agg formulation and torch.sparse.mm at cfg5 size
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import make_products_graph, SPMM_N, SPMM_F
    rowptr, col, val, _ = make_products_graph(DEV)
    x = torch.randn(SPMM_N, SPMM_F, device=DEV)
    dst = torch.repeat_interleave(torch.arange(SPMM_N, device=DEV), rowptr[1:] - rowptr[:-1])
    src = col.long()

    def ref_agg():
        out = torch.zeros(SPMM_N, SPMM_F, device=DEV)
        return out.scatter_add_(0, dst[:, None].expand(-1, SPMM_F), x[src] * val[:, None])
    ms = timeit(ref_agg, 3, 1)
    emit(section="torch_gpu", what="reference edge _agg (index * w -> scatter_add_) on stock torch CUDA, cfg5", ms=ms,
         edges_per_s=col.numel() / ms * 1e3)
    A = torch.sparse_csr_tensor(rowptr, col.long(), val, size=(SPMM_N, SPMM_N))
    ms_s = timeit(lambda: torch.sparse.mm(A, x), 3, 1)
    emit(section="torch_gpu", what="torch.sparse.mm CSR (cuSPARSE), cfg5", ms=ms_s, edges_per_s=col.numel() / ms_s * 1e3)
    ms_o = timeit(lambda: ops.csr_spmm(rowptr, col, val, x), 5, 2)
    emit(section="torch_gpu", what="ours csr_spmm, cfg5", ms=ms_o, edges_per_s=col.numel() / ms_o * 1e3,
         speedup_vs_scatter=ms / ms_o, speedup_vs_cusparse=ms_s / ms_o)


class _PM:
    def __init__(self, emb): self.emb = emb
    def inference(self, features, adj): return self.emb


def small():
    """cfg1 (Cora-shaped node forward) and cfg2 (graph variant, Q=1) latency vs the reference formulas on stock torch."""
    g = torch.Generator(device=DEV).manual_seed(0)
    n, d, C, N = 2708, 256, 3, 10832
    a = (torch.rand(n, n, generator=g, device=DEV) < 3.9 / n).float()
    a = torch.triu(a, 1); a = a + a.t() + torch.eye(n, device=DEV)
    dinv = a.sum(1).pow(-0.5); adj = dinv[:, None] * a * dinv[None, :]
    emb = torch.randn(n, d, generator=g, device=DEV)
    keys = F.normalize(torch.randn(N, d, generator=g, device=DEV), dim=-1); vals = torch.randn(N, d, generator=g, device=DEV)
    labs = F.one_hot(torch.randint(0, C, (N,), generator=g, device=DEV), C).float()
    base = R.ToyGraphBase(None, C, d, 3); base.add_entries(keys, vals, labs)
    model = R.RAGraph(_PM(emb), base, 0, C, d).to(DEV).eval()
    dec = model.decoder

    def ours():
        with torch.no_grad():
            return model(None, adj)

    def ref():
        with torch.no_grad():
            s = torch.matmul(F.normalize(emb, dim=-1), F.normalize(keys, dim=-1).t())
            _, i = torch.topk(s, C + 1)
            re, rl = vals[i].sum(1), labs[i].mean(1)
            deg = adj.sum(1, keepdim=True); an = adj / deg; x = emb
            for _ in range(3):
                x = F.relu(an @ x)
            h = x * 0.5 + re * 0.5
            return torch.softmax(dec(h), 1) * 0.5 + rl * 0.5
    o, r = ours(), ref()
    emit(section="small", cfg="cfg1 Cora-shaped node forward (n=2708, d=256, N=10832, k=4, 3 hops)", ours_ms=timeit(ours, 20),
         stock_torch_ms=timeit(ref, 20), max_abs_diff=float((o - r).abs().max()))
    # cfg2: one query vector against a few hundred graph keys
    Ng = 480
    gb = R.ToyGraphBase(None, 6, d, 1, variant="graph")
    gk = torch.randn(Ng, d, generator=g, device=DEV) * 0.3; gv = torch.randn(Ng, d, generator=g, device=DEV)
    gl = F.one_hot(torch.randint(0, 6, (Ng,), generator=g, device=DEV), 6)
    gb.add_entries(gk, gv, gl)
    q1 = torch.randn(d, generator=g, device=DEV)

    def ours2():
        return gb.retrieve(q1, None, False)

    def ref2():
        s = torch.matmul(F.normalize(q1, dim=-1), F.normalize(gk, dim=-1).t()).unsqueeze(0)
        _, i = torch.topk(s, 3)
        return gv[i], gl[i]
    e1, e2 = ours2()[0], ref2()[0]
    emit(section="small", cfg="cfg2 graph variant retrieve (Q=1, N=480, d=256, k=3)", ours_us=timeit(ours2, 50) * 1e3,
         stock_torch_us=timeit(ref2, 50) * 1e3, equal=bool(torch.equal(e1, e2)))


def spmm_variants():
    from bench import make_products_graph, SPMM_N
    rowptr, col, val, md = make_products_graph(DEV)
    for Fdim in (64, 128, 256):
        x = torch.randn(SPMM_N, Fdim, device=DEV)
        ms = timeit(lambda: ops.csr_spmm(rowptr, col, val, x), 5, 2)
        alg = col.numel() * 8 + (SPMM_N + 1) * 8 + col.numel() * Fdim * 4 + SPMM_N * Fdim * 4
        emit(section="spmm_variants", F=Fdim, ms=ms, gbs=alg / ms / 1e6, frac_of_hbm=alg / ms / 1e6 / PEAKS["hbm_gbs"],
             edges_per_s=col.numel() / ms * 1e3)
    x = torch.randn(SPMM_N, 256, device=DEV)
    epi = L.EPI_ROWNORM | L.EPI_RELU
    ms = timeit(lambda: ops.csr_spmm(rowptr, col, val, x, epi), 5, 2)
    emit(section="spmm_variants", F=256, epilogue="ROWNORM|RELU (Propagation hop)", ms=ms)
    # uniform-random graph of the same size (no hubs -> little L2 reuse)
    n, nnz = SPMM_N, col.numel()
    g = torch.Generator(device=DEV).manual_seed(3)
    dst = torch.sort(torch.randint(0, n, (nnz,), generator=g, device=DEV)).values
    rp = torch.zeros(n + 1, dtype=torch.int64, device=DEV); torch.cumsum(torch.bincount(dst, minlength=n), 0, out=rp[1:])
    cu = torch.randint(0, n, (nnz,), generator=g, device=DEV, dtype=torch.int32)
    ms = timeit(lambda: ops.csr_spmm(rp, cu, val, x), 5, 2)
    alg = nnz * 8 + (n + 1) * 8 + nnz * 256 * 4 + n * 256 * 4
    emit(section="spmm_variants", graph="uniform random (same n, nnz)", F=256, ms=ms, gbs=alg / ms / 1e6,
         frac_of_hbm=alg / ms / 1e6 / PEAKS["hbm_gbs"])


if __name__ == "__main__":
    todo = sys.argv[1:] or ["cfg3", "gather", "torch_gpu", "small", "spmm_variants"]
    for name in todo:
        try:
            globals()[name]()
        except Exception as e:          # keep going: each section is independent
            emit(section=name, error=repr(e)[:300])
        torch.cuda.empty_cache()
