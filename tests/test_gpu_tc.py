"""GPU parity tests of the tcgen05 tensor-core retrieval modes (bf16 filter + fp32 refine = exact; bf16 raw)."""
import numpy as np
import pytest
import torch

from ragraph_b200 import _lib as L
from ragraph_b200 import ops
from oracle import ragraph_oracle as O

import os

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = "cuda"


@pytest.fixture(params=["ts", "ss", "auto"], autouse=True)
def tc_variant(request, monkeypatch):
    """ts = query tile stationary in tensor memory (long key streams); ss = query tile in shared memory (short streams);
    auto = the library's own choice by stream length."""
    if request.param == "auto":
        monkeypatch.delenv("RAG_TC_VARIANT", raising=False)
    else:
        monkeypatch.setenv("RAG_TC_VARIANT", request.param)
    return request.param


def _run(q, keys, k, mode):
    qd, kd = q.to(DEV), keys.to(DEV)
    inv = ops.row_inv_norm(kd)
    shadow = ops.rows_to_bf16(kd, True)
    s, i = ops.cosine_topk(qd, kd, k, key_inv_norm=inv, keys_bf16=shadow, mode=mode)
    torch.cuda.synchronize()
    return s.cpu(), i.cpu()


@pytest.mark.parametrize("Q,N,d,k", [(300, 20000, 128, 10), (700, 100000, 64, 10), (1000, 50000, 256, 10),
                                     (37, 5000, 128, 20), (256, 128, 128, 4), (5, 333, 100, 3), (513, 70001, 128, 10),
                                     (64, 3000, 32, 26), (260, 30000, 160, 10), (300, 40000, 256, 7),
                                     (1100, 60000, 128, 10)])
def test_refine_mode_is_exact(Q, N, d, k):
    g = torch.Generator().manual_seed(Q + N + d)
    q = torch.randn(Q, d, generator=g); keys = torch.randn(N, d, generator=g)
    keys[N // 2] = keys[1]                      # exact duplicate -> tie
    keys[3] = 0.0; q[Q // 2] = 0.0              # zero rows: eps clamp
    assert L.load().rag_sim_mode_supported(L.SIM_BF16_REFINE, d, k)
    s3, i3 = _run(q, keys, k, L.SIM_BF16_REFINE)
    S64 = O.cosine_similarity_f64(q.numpy(), keys.numpy())
    ok, bad = O.topk_sets_match(i3.numpy(), S64, k)
    assert ok, bad[:5]
    exact = np.take_along_axis(S64, i3.numpy(), axis=1)
    assert np.max(np.abs(s3.numpy() - exact)) < 1e-5
    assert np.all(s3.numpy()[:, :-1] >= s3.numpy()[:, 1:])
    s0, i0 = ops.cosine_topk(q.to(DEV), keys.to(DEV), k)             # fp32 CUDA-core path
    assert np.max(np.abs(s0.cpu().numpy() - s3.numpy())) < 2e-6
    same = (i0.cpu() == i3).all(dim=1)
    rows = torch.nonzero(~same).flatten().tolist()                  # any difference must be a tie within 1e-6
    for r in rows:
        a = np.sort(S64[r, i0[r].cpu().numpy()]); b = np.sort(S64[r, i3[r].numpy()])
        assert np.max(np.abs(a - b)) < 1e-6, r


def test_refine_mode_clustered_keys_fall_back_to_fp32():
    """Near-duplicate keys (sigma 1e-3 around few centroids, like Augmentation.augment_features at tiny sigma) defeat
    the bf16 certificate; those rows must be recomputed exactly by the fp32 kernel."""
    g = torch.Generator().manual_seed(9)
    d, N, Q, k = 128, 40000, 300, 10
    cent = torch.randn(20, d, generator=g)
    keys = cent[torch.randint(0, 20, (N,), generator=g)] + 1e-3 * torch.randn(N, d, generator=g)
    q = cent[torch.randint(0, 20, (Q,), generator=g)] + 0.05 * torch.randn(Q, d, generator=g)
    s3, i3 = _run(q, keys, k, L.SIM_BF16_REFINE)
    S64 = O.cosine_similarity_f64(q.numpy(), keys.numpy())
    ok, bad = O.topk_sets_match(i3.numpy(), S64, k)
    assert ok, bad[:5]
    assert np.max(np.abs(s3.numpy() - np.take_along_axis(S64, i3.numpy(), axis=1))) < 1e-5


def test_bf16_raw_mode_recall():
    g = torch.Generator().manual_seed(21)
    Q, N, d, k = 512, 200000, 128, 10
    q = torch.randn(Q, d, generator=g); keys = torch.randn(N, d, generator=g)
    s2, i2 = _run(q, keys, k, L.SIM_BF16)
    ref_s, ref_i = O.topk(O.cosine_similarity(q, keys), k)
    rec = O.recall_at_k(i2.numpy(), ref_i.numpy())
    assert rec > 0.9, rec
    assert float((s2 - ref_s).abs().max()) < 5e-3                   # bf16 products: |err| <= 2^-8


def test_tc_unsupported_shapes_raise():
    q, keys = torch.randn(8, 320, device=DEV), torch.randn(1000, 320, device=DEV)
    assert not L.load().rag_sim_mode_supported(L.SIM_BF16_REFINE, 320, 10)
    with pytest.raises(L.RagError, match="RAG_EUNSUPPORTED"):
        ops.cosine_topk(q, keys, 10, key_inv_norm=ops.row_inv_norm(keys), keys_bf16=ops.rows_to_bf16(keys), mode=3)
    with pytest.raises(L.RagError, match="RAG_EINVAL"):
        ops.cosine_topk(q[:, :128].contiguous(), keys[:, :128].contiguous(), 10, mode=3)


def test_large_refine_equals_fp32():
    torch.manual_seed(1)
    N, d, Q, k = 2_000_000, 128, 1024, 10
    keys = torch.nn.functional.normalize(torch.randn(N, d, device=DEV), dim=-1); q = torch.randn(Q, d, device=DEV)
    inv = ops.row_inv_norm(keys); shadow = ops.rows_to_bf16(keys, True)
    s3, i3 = ops.cosine_topk(q, keys, k, key_inv_norm=inv, keys_bf16=shadow, mode=L.SIM_BF16_REFINE)
    s0, i0 = ops.cosine_topk(q, keys, k, key_inv_norm=inv)
    assert float((s3 - s0).abs().max()) < 2e-6
    assert float((i3 == i0).float().mean()) > 0.9999


@pytest.mark.parametrize("min_tiles", [64, 1024])
def test_prepass_threshold_keeps_exactness(min_tiles, monkeypatch, tc_variant):
    """The threshold pre-pass (group maxima over 1/64 of every split -> per-row lower bound of the k'-th best score)
    must not change results: duplicates of the best match (ties AT the bound), zero rows, and queries whose whole top-k
    sits inside the sampled prefix."""
    if tc_variant != "ts":
        pytest.skip("pre-pass exists in the ts kernel only")
    monkeypatch.setenv("RAG_TC_PREPASS_MIN_TILES", str(min_tiles))
    torch.manual_seed(11)
    Q, d, k = 4096, 128, 10
    N = 1_300_000 if min_tiles == 1024 else 120_000
    keys = torch.randn(N, d, device=DEV); q = torch.randn(Q, d, device=DEV)
    q[:64] = keys[:64] + 0.01 * torch.randn(64, d, device=DEV)         # best matches inside the sampled prefix of split 0
    keys[1000:1020] = keys[5]                                          # 20 exact duplicates of a key that query 5 matches
    keys[N - 30:N - 10] = q[100] * 3.0                                 # 20 identical best matches for query 100 (cos = 1)
    q[7] = 0.0; keys[9] = 0.0
    inv = ops.row_inv_norm(keys); shadow = ops.rows_to_bf16(keys, True)
    s3, i3 = ops.cosine_topk(q, keys, k, key_inv_norm=inv, keys_bf16=shadow, mode=L.SIM_BF16_REFINE)
    monkeypatch.setenv("RAG_TC_PREPASS", "0")
    s3n, i3n = ops.cosine_topk(q, keys, k, key_inv_norm=inv, keys_bf16=shadow, mode=L.SIM_BF16_REFINE)
    s0, i0 = ops.cosine_topk(q, keys, k, key_inv_norm=inv)             # fp32 CUDA-core path
    assert torch.equal(s3, s3n) and torch.equal(i3, i3n)
    assert float((s3 - s0).abs().max()) < 2e-6
    same = (i3 == i0).all(dim=1)
    assert float(same.float().mean()) > 0.99
    for r in torch.nonzero(~same).flatten().tolist():                  # any difference must be a tie within 1e-6
        assert float((s3[r].sort().values - s0[r].sort().values).abs().max()) < 1e-6
