"""GPU parity tests of the tcgen05 tensor-core retrieval modes: 16-bit filter (fp16 = the library default, bf16) + fp32
refine = exact; raw 16-bit modes by recall; the second tensor-core pass for rows the first cannot certify."""
import numpy as np
import pytest
import torch

from ragraph_b200 import _lib as L
from ragraph_b200 import ops
from oracle import ragraph_oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = "cuda"
EXACT_MODES = [L.SIM_F16_REFINE, L.SIM_BF16_REFINE]
FMT = {L.SIM_F16_REFINE: L.FMT_F16, L.SIM_F16: L.FMT_F16, L.SIM_BF16_REFINE: L.FMT_BF16, L.SIM_BF16: L.FMT_BF16}


@pytest.fixture(params=["ts", "ss", "auto"], autouse=True)
def tc_variant(request):
    """ts = query tile stationary in tensor memory (long key streams); ss = query tile in shared memory (short streams);
    auto = the library's own choice by stream length."""
    L.tc_set_option("variant", {"auto": 0, "ss": 1, "ts": 2}[request.param])
    yield request.param
    for name in ("variant", "prepass", "prepass_min_tiles", "kp", "pass2", "gshare"):
        L.tc_set_option(name, -1)


def _shadow(kd, mode, with_err=True):
    err = torch.zeros(1, device=kd.device) if with_err else None
    sh, _ = ops.rows_to_shadow16(kd, FMT[mode], True, err_max=err)
    return sh, err


def _run(q, keys, k, mode, stats=False, with_err=True):
    qd, kd = q.to(DEV), keys.to(DEV)
    inv = ops.row_inv_norm(kd)
    shadow, err = _shadow(kd, mode, with_err)
    out = ops.cosine_topk_with_stats(qd, kd, k, inv, shadow, mode, shadow_err=err)
    torch.cuda.synchronize()
    return (out[0].cpu(), out[1].cpu(), out[2].cpu().tolist()) if stats else (out[0].cpu(), out[1].cpu())


def _assert_exact(q, keys, k, s3, i3):
    S64 = O.cosine_similarity_f64(q.numpy(), keys.numpy())
    ok, bad = O.topk_sets_match(i3.numpy(), S64, k)
    assert ok, bad[:5]
    exact = np.take_along_axis(S64, i3.numpy(), axis=1)
    assert np.max(np.abs(s3.numpy() - exact)) < 1e-5
    assert np.all(s3.numpy()[:, :-1] >= s3.numpy()[:, 1:])
    return S64


@pytest.mark.parametrize("mode", EXACT_MODES)
@pytest.mark.parametrize("Q,N,d,k", [(300, 20000, 128, 10), (700, 100000, 64, 10), (1000, 50000, 256, 10),
                                     (37, 5000, 128, 20), (256, 128, 128, 4), (5, 333, 100, 3), (513, 70001, 128, 10),
                                     (64, 3000, 32, 26), (260, 30000, 160, 10), (300, 40000, 256, 7),
                                     (1100, 60000, 128, 10),
                                     # k in (26, 128]: more key splits instead of longer lists (edge vanilla phase: retrieve_num = 50)
                                     (300, 40000, 64, 50), (70, 30000, 128, 128), (600, 90000, 64, 51), (33, 2000, 96, 100)])
def test_refine_mode_is_exact(Q, N, d, k, mode):
    g = torch.Generator().manual_seed(Q + N + d)
    q = torch.randn(Q, d, generator=g); keys = torch.randn(N, d, generator=g)
    keys[N // 2] = keys[1]                      # exact duplicate -> tie
    keys[3] = 0.0; q[Q // 2] = 0.0              # zero rows: eps clamp
    assert L.load().rag_sim_mode_supported(mode, d, k)
    s3, i3 = _run(q, keys, k, mode)
    S64 = _assert_exact(q, keys, k, s3, i3)
    s0, i0 = ops.cosine_topk(q.to(DEV), keys.to(DEV), k)             # fp32 CUDA-core path
    assert np.max(np.abs(s0.cpu().numpy() - s3.numpy())) < 2e-6
    same = (i0.cpu() == i3).all(dim=1)
    rows = torch.nonzero(~same).flatten().tolist()                  # any difference must be a tie within 1e-6
    for r in rows:
        a = np.sort(S64[r, i0[r].cpu().numpy()]); b = np.sort(S64[r, i3[r].numpy()])
        assert np.max(np.abs(a - b)) < 1e-6, r


def _clustered(g, N, Q, d, n_cent, sigma_k, sigma_q):
    cent = torch.randn(n_cent, d, generator=g)
    keys = cent[torch.randint(0, n_cent, (N,), generator=g)] + sigma_k * torch.randn(N, d, generator=g)
    q = cent[torch.randint(0, n_cent, (Q,), generator=g)] + sigma_q * torch.randn(Q, d, generator=g)
    return q, keys


@pytest.mark.parametrize("mode", EXACT_MODES)
@pytest.mark.parametrize("d,k", [(128, 10), (256, 10), (64, 20), (64, 50)])
def test_clustered_library_second_pass_is_exact(mode, d, k):
    """Realistic library (Gaussian centroids, sigma = 0.1 like Augmentation.augment_features, Augmentation.py:8-20):
    thousands of keys score within the 16-bit error bound of the k-th best, the first pass cannot certify those rows, the
    second tensor-core pass (collect everything above exact k-th - eps) must make them exact -- without the fp32 kernel."""
    g = torch.Generator().manual_seed(9 + d)
    q, keys = _clustered(g, 60000, 300, d, 8, 0.1, 0.1)
    s3, i3, st = _run(q, keys, k, mode, stats=True)
    _assert_exact(q, keys, k, s3, i3)
    if mode == L.SIM_F16_REFINE:
        if k <= 26:                         # (k = 50: more than 1 024 keys within the bound of the 50th best may overflow)
            assert st[1] == 0, f"rows fell through to the fp32 kernel: {st}"
    else:                                   # bf16: ~4 sigma of the in-cluster score spread -> thousands of near-ties per row
        assert st[0] > 0, "the bf16 certificate cannot hold on this library: the second pass must have run"


@pytest.mark.parametrize("mode", EXACT_MODES)
def test_near_duplicate_clusters_overflow_to_fp32(mode):
    """sigma = 1e-3 around 20 centroids: ~2000 keys within any 16-bit bound of each row's k-th best -- more than the second
    pass keeps per row (1024), so those rows must end in the fp32 kernel and still be exact."""
    g = torch.Generator().manual_seed(9)
    q, keys = _clustered(g, 40000, 300, 128, 20, 1e-3, 0.05)
    s3, i3, st = _run(q, keys, 10, mode, stats=True)
    _assert_exact(q, keys, 10, s3, i3)
    assert st[0] > 0 and st[1] > 0, st


@pytest.mark.parametrize("mode", EXACT_MODES)
def test_duplicate_rows_5pct(mode):
    """5 % exact duplicate rows (the reference's multinomial(replacement=True) draw, ToyGraphBase.py:98): ties inside the
    top k; index order must be the deterministic one (score desc, index asc) and equal the fp32 kernel's."""
    g = torch.Generator().manual_seed(77)
    Q, N, d, k = 400, 50000, 128, 10
    q = torch.randn(Q, d, generator=g); keys = torch.randn(N, d, generator=g)
    dup = torch.randperm(N, generator=g)[:N // 20]
    keys[dup] = keys[(dup + 7919) % N]
    q[:100] = keys[dup[:100]] + 0.05 * torch.randn(100, d, generator=g)    # queries whose best match IS a duplicated row
    s3, i3 = _run(q, keys, k, mode)
    _assert_exact(q, keys, k, s3, i3)
    s0, i0 = ops.cosine_topk(q.to(DEV), keys.to(DEV), k)
    agree = float((i0.cpu() == i3).float().mean())
    assert agree > 0.999, agree


def test_second_pass_off_matches_on():
    """pass2 = 0 (uncertified rows straight to the fp32 kernel, the round-1 behaviour) and pass2 = 1 return the same."""
    g = torch.Generator().manual_seed(5)
    q, keys = _clustered(g, 50000, 260, 128, 6, 0.1, 0.1)
    s1, i1, st1 = _run(q, keys, 10, L.SIM_BF16_REFINE, stats=True)
    L.tc_set_option("pass2", 0)
    s0, i0, st0 = _run(q, keys, 10, L.SIM_BF16_REFINE, stats=True)
    assert st1[0] > 0 and st0[0] == st1[0]
    assert float((s0 - s1).abs().max()) < 2e-6
    assert float((i0 == i1).float().mean()) > 0.995                 # near-ties may swap: both are checked against fp64
    _assert_exact(q, keys, 10, s0, i0)
    _assert_exact(q, keys, 10, s1, i1)


def test_certificate_without_measured_error_norm():
    """shadow_err = None: the certificate uses the element-wise worst case (2 u); results are the same, more rows take the
    second pass."""
    g = torch.Generator().manual_seed(6)
    q = torch.randn(500, 128, generator=g); keys = torch.randn(80000, 128, generator=g)
    sa, ia, sta = _run(q, keys, 10, L.SIM_F16_REFINE, stats=True, with_err=True)
    sb, ib, stb = _run(q, keys, 10, L.SIM_F16_REFINE, stats=True, with_err=False)
    assert float((sa - sb).abs().max()) < 2e-6 and float((ia == ib).float().mean()) > 0.999
    assert stb[0] >= sta[0]


def test_fp16_subnormal_components():
    """Key components below the fp16 normal range (|x| < 2^-14 after normalisation) are flushed in the shadow and the flush
    is part of the measured error norm -- rows dominated by tiny components must still come out exact."""
    g = torch.Generator().manual_seed(8)
    Q, N, d, k = 256, 30000, 128, 10
    keys = torch.randn(N, d, generator=g); q = torch.randn(Q, d, generator=g)
    keys[:, 64:] *= 3e-5                                             # half of every key sits in the subnormal range
    q[:, 64:] *= 50.0                                                # ... and the queries weigh exactly that half
    s3, i3 = _run(q, keys, k, L.SIM_F16_REFINE)
    _assert_exact(q, keys, k, s3, i3)


@pytest.mark.parametrize("mode,tol,rec_min", [(L.SIM_BF16, 5e-3, 0.9), (L.SIM_F16, 7e-4, 0.985)])
def test_raw_16bit_mode_recall(mode, tol, rec_min):
    g = torch.Generator().manual_seed(21)
    Q, N, d, k = 512, 200000, 128, 10
    q = torch.randn(Q, d, generator=g); keys = torch.randn(N, d, generator=g)
    s2, i2 = _run(q, keys, k, mode)
    ref_s, ref_i = O.topk(O.cosine_similarity(q, keys), k)
    rec = O.recall_at_k(i2.numpy(), ref_i.numpy())
    assert rec > rec_min, rec
    assert float((s2 - ref_s).abs().max()) < tol                    # |err| <= 2 u: bf16 2^-7, fp16 2^-10 (worst case)


def test_tc_unsupported_shapes_raise():
    q, keys = torch.randn(8, 320, device=DEV), torch.randn(1000, 320, device=DEV)
    assert not L.load().rag_sim_mode_supported(L.SIM_BF16_REFINE, 320, 10)
    assert not L.load().rag_sim_mode_supported(L.SIM_F16_REFINE, 320, 10)
    with pytest.raises(L.RagError, match="RAG_EUNSUPPORTED"):
        ops.cosine_topk(q, keys, 10, key_inv_norm=ops.row_inv_norm(keys), keys_bf16=ops.rows_to_bf16(keys), mode=3)
    with pytest.raises(L.RagError, match="RAG_EINVAL"):
        ops.cosine_topk(q[:, :128].contiguous(), keys[:, :128].contiguous(), 10, mode=3)
    with pytest.raises(RuntimeError, match="shadow"):                 # bf16 shadow handed to an fp16 mode
        k128 = keys[:, :128].contiguous()
        ops.cosine_topk(q[:, :128].contiguous(), k128, 10, key_inv_norm=ops.row_inv_norm(k128),
                        keys_bf16=ops.rows_to_bf16(k128), mode=L.SIM_F16_REFINE)


@pytest.mark.parametrize("mode", EXACT_MODES)
def test_large_refine_equals_fp32(mode):
    torch.manual_seed(1)
    N, d, Q, k = 2_000_000, 128, 1024, 10
    keys = torch.nn.functional.normalize(torch.randn(N, d, device=DEV), dim=-1); q = torch.randn(Q, d, device=DEV)
    inv = ops.row_inv_norm(keys); shadow, err = _shadow(keys, mode)
    s3, i3 = ops.cosine_topk(q, keys, k, key_inv_norm=inv, keys_bf16=shadow, mode=mode, shadow_err=err)
    s0, i0 = ops.cosine_topk(q, keys, k, key_inv_norm=inv)
    assert float((s3 - s0).abs().max()) < 2e-6
    same = (i3 == i0).all(dim=1)
    rows = torch.nonzero(~same).flatten().tolist()
    assert len(rows) <= 2, len(rows)
    for r in rows:                                                    # tie-aware: fp64 scores of the union of both id sets
        ids = torch.unique(torch.cat([i3[r], i0[r]]))
        s64 = (torch.nn.functional.normalize(q[r].double(), dim=-1)[None] * torch.nn.functional.normalize(keys[ids].double(), dim=-1)).sum(-1)
        a = s64[torch.isin(ids, i3[r])].sort().values; b = s64[torch.isin(ids, i0[r])].sort().values
        assert float((a - b).abs().max()) < 1e-6, r


@pytest.mark.parametrize("kp", [16, 32])
def test_list_length_option(kp, tc_variant):
    """rag_tc_set_option("kp", 32): 32-entry candidate lists for k <= 10 too (d <= 128) -- same exact results."""
    L.tc_set_option("kp", kp)
    g = torch.Generator().manual_seed(3)
    q, keys = _clustered(g, 60000, 300, 128, 8, 0.1, 0.1)
    s3, i3 = _run(q, keys, 10, L.SIM_F16_REFINE)
    _assert_exact(q, keys, 10, s3, i3)


@pytest.mark.parametrize("min_tiles", [64, 1024])
def test_prepass_threshold_keeps_exactness(min_tiles, tc_variant):
    """The threshold pre-pass (group maxima over 1/64 of every split -> per-row lower bound of the k'-th best score)
    must not change results: duplicates of the best match (ties AT the bound), zero rows, and queries whose whole top-k
    sits inside the sampled prefix."""
    if tc_variant != "ts":
        pytest.skip("pre-pass exists in the ts kernel only")
    L.tc_set_option("prepass_min_tiles", min_tiles)
    torch.manual_seed(11)
    Q, d, k = 4096, 128, 10
    N = 1_300_000 if min_tiles == 1024 else 120_000
    keys = torch.randn(N, d, device=DEV); q = torch.randn(Q, d, device=DEV)
    q[:64] = keys[:64] + 0.01 * torch.randn(64, d, device=DEV)         # best matches inside the sampled prefix of split 0
    keys[1000:1020] = keys[5]                                          # 20 exact duplicates of a key that query 5 matches
    keys[N - 30:N - 10] = q[100] * 3.0                                 # 20 identical best matches for query 100 (cos = 1)
    q[7] = 0.0; keys[9] = 0.0
    inv = ops.row_inv_norm(keys)
    for mode in EXACT_MODES:
        shadow, err = _shadow(keys, mode)
        L.tc_set_option("prepass", 1)
        s3, i3 = ops.cosine_topk(q, keys, k, key_inv_norm=inv, keys_bf16=shadow, mode=mode, shadow_err=err)
        L.tc_set_option("prepass", 0)
        s3n, i3n = ops.cosine_topk(q, keys, k, key_inv_norm=inv, keys_bf16=shadow, mode=mode, shadow_err=err)
        s0, i0 = ops.cosine_topk(q, keys, k, key_inv_norm=inv)             # fp32 CUDA-core path
        assert torch.equal(s3, s3n) and torch.equal(i3, i3n)
        assert float((s3 - s0).abs().max()) < 2e-6
        same = (i3 == i0).all(dim=1)
        assert float(same.float().mean()) > 0.99
        for r in torch.nonzero(~same).flatten().tolist():                  # any difference must be a tie within 1e-6
            assert float((s3[r].sort().values - s0[r].sort().values).abs().max()) < 1e-6


@pytest.mark.parametrize("mode", EXACT_MODES + [L.SIM_BF16])
def test_cross_split_threshold_sharing_same_results(mode):
    """The sweeping CTAs (cross-split threshold sharing, TcArgs::pool) only ever raise a row's threshold to a proven lower
    bound of what refine needs, so the answers with and without them are bit-identical -- and with them the filter queues
    several times fewer hits.  Shape: 10 query tiles x 14 key splits = 140 worker CTAs (idle SMs left), >= 1024 tiles each."""
    torch.manual_seed(77)
    Q, N, d, k = 2400, 1_900_000, 64, 10
    kd = torch.randn(N, d, device=DEV)
    kd[N // 3] = kd[5]                                   # an exact duplicate: a tie inside the top k of the rows near it
    qd = torch.randn(Q, d, device=DEV)
    qd[7] = kd[5] + 0.01 * torch.randn(d, device=DEV)
    inv = ops.row_inv_norm(kd)
    shadow, err = _shadow(kd, mode)
    L.tc_set_option("variant", 2)
    res = {}
    for g in (0, 1):
        L.tc_set_option("gshare", g)
        s, i, st = ops.cosine_topk_with_stats(qd, kd, k, inv, shadow, mode, shadow_err=err)
        res[g] = (s.clone(), i.clone(), st.tolist())
    L.tc_set_option("gshare", -1); L.tc_set_option("variant", -1)
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    if mode != L.SIM_BF16:                                                         # (the raw modes leave no counters)
        assert res[0][2][3] == 0 and res[1][2][3] == 140, (res[0][2], res[1][2])  # worker CTAs seen by the sweep
        assert res[1][2][4] < res[0][2][4], (res[0][2], res[1][2])                # fewer hits queued
        rows = torch.arange(0, Q, 16, device=DEV)
        s0, i0 = ops.cosine_topk(qd[rows].contiguous(), kd, k, inv)
        assert float((res[1][0][rows] - s0).abs().max()) < 2e-6
        diff = (res[1][1][rows] != i0).any(dim=1)
        for r in torch.nonzero(diff).flatten().tolist():                           # any difference must be a tie
            assert float((res[1][0][rows][r] - s0[r]).abs().max()) < 1e-6


@pytest.mark.parametrize("mode", EXACT_MODES)
@pytest.mark.parametrize("kind", ["gauss", "clustered", "tiny_clusters"])
def test_two_pass_mode_same_results(mode, kind):
    """Short key streams (automatic kernel choice): a maxima pass + a collect pass above (k-th largest group maximum - 2 eps)
    replace the list warm-up.  Same answers as the one-pass path and as fp64; rows whose collect area overflows (a bf16 filter
    on a clustered library) are retried with a threshold from their exact scores (second pass) before the fp32 kernel."""
    g = torch.Generator().manual_seed(31)
    Q, N, d, k = 600, 70000, 128, 10
    if kind == "gauss":
        q, keys = torch.randn(Q, d, generator=g), torch.randn(N, d, generator=g)
        keys[N // 2] = keys[1]; keys[3] = 0.0; q[Q // 2] = 0.0
    elif kind == "clustered":
        q, keys = _clustered(g, N, Q, d, 8, 0.1, 0.1)             # ~9 000 keys per cluster: the collect areas (1 024) overflow
    else:
        q, keys = _clustered(g, N, Q, d, 2000, 0.1, 0.1)          # ~35 near neighbours per query: the reference's kind of library
    L.tc_set_option("variant", 0)
    out = {}
    for tp in (0, 1):
        L.tc_set_option("twopass", tp)
        out[tp] = _run(q, keys, k, mode, stats=True)
    L.tc_set_option("twopass", -1)
    _assert_exact(q, keys, k, out[1][0], out[1][1])
    assert float((out[0][0] - out[1][0]).abs().max()) < 2e-6
    assert float((out[0][1] == out[1][1]).all(dim=1).float().mean()) > 0.99          # near-ties may swap; both checked against fp64
    if kind == "clustered":
        if mode == L.SIM_F16_REFINE:        # fp16: a few hundred keys above the bound per row -- no row needs the fp32 kernel
            assert out[1][2][1] == 0, out[1][2]
        else:                               # bf16: the whole cluster lies inside the bound -> collect areas overflow -> retry list
            assert out[1][2][0] > 0, out[1][2]
    if kind == "tiny_clusters":
        assert out[1][2][:2] == [0, 0], out[1][2]
