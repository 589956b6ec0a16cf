"""RAG_SIM_TF32: the tcgen05 kind::tf32 filter (raw approximate scores), reported by recall@k and score error like the
north star asks for the tensor-core modes.  This file sorts last on purpose: it is the newest kernel path."""
import numpy as np
import pytest
import torch

import ragraph_b200 as R
from ragraph_b200 import _lib as L
from ragraph_b200 import ops
from oracle import ragraph_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
TF32_ERR = 2.0 ** -10 + 1e-5          # two operands rounded to 11 significant bits (unit vectors) + fp32 accumulation


def _case(Q, N, d, seed, clustered=False):
    g = torch.Generator().manual_seed(seed)
    keys = torch.randn(N, d, generator=g)
    if clustered:                                   # near-duplicate keys as Augmentation.augment_features makes them
        cent = torch.randn(64, d, generator=g)
        keys = cent[torch.randint(0, 64, (N,), generator=g)] + 0.1 * keys
    q = torch.randn(Q, d, generator=g)
    q[1] = 0.0
    return q, keys


@pytest.mark.parametrize("Q,N,d,k", [(300, 70000, 128, 10), (257, 40000, 64, 26), (64, 9000, 32, 4),
                                     (130, 33333, 100, 10), (33, 20000, 48, 16), (5, 700, 16, 1)])
def test_tf32_topk_recall_and_error(Q, N, d, k):
    if not L.load().rag_sim_mode_supported(L.SIM_TF32, d, k):
        pytest.skip("tf32 mode not built for this shape")
    q, keys = _case(Q, N, d, 1000 + d + k)
    qd, kd = q.to(DEV), keys.to(DEV)
    shadow = ops.rows_to_tf32(kd)
    assert shadow.shape == (N, ops.tf32_shadow_dpad(d)) and shadow.dtype == torch.float32
    # the shadow is the normalised key matrix rounded to tf32: low 13 mantissa bits are zero
    assert int((shadow.view(torch.int32) & 0x1fff).abs().max()) == 0
    ref_n = torch.nn.functional.normalize(keys, dim=-1)
    assert float((shadow[:, :d].cpu() - ref_n).abs().max()) < 2.0 ** -11 + 1e-7
    assert bool((shadow[:, d:] == 0).all())
    s, i = ops.cosine_topk(qd, kd, k, ops.row_inv_norm(kd), shadow, L.SIM_TF32)
    s, i = s.cpu(), i.cpu()
    S64 = O.cosine_similarity_f64(q, keys)
    assert bool(((i >= 0) & (i < N)).all()) and bool((s[:, :-1] >= s[:, 1:]).all())
    for r in range(Q):
        assert len(set(i[r].tolist())) == k
    exact_of_returned = np.take_along_axis(S64, i.numpy(), axis=1)
    assert float(np.abs(s.numpy().astype(np.float64) - exact_of_returned).max()) < TF32_ERR
    _, ref_i = O.topk(torch.from_numpy(S64), k)
    rows = [r for r in range(Q) if r != 1]           # row 1 is the zero query: every key ties at score 0
    recall = O.recall_at_k(i.numpy()[rows], ref_i.numpy()[rows])
    print(f"tf32 recall@{k} (Q={Q}, N={N}, d={d}): {recall:.4f}")
    assert recall >= 0.98, recall
    # every returned key is within the tf32 error band of the exact k-th best: nothing far below the cut is returned
    kth = np.sort(S64, axis=1)[:, -k]
    assert bool((exact_of_returned >= kth[:, None] - 2 * TF32_ERR).all())


def test_tf32_store_mode_and_clustered_recall():
    d, k = 64, 10
    if not L.load().rag_sim_mode_supported(L.SIM_TF32, d, k):
        pytest.skip("tf32 mode not built for this shape")
    q, keys = _case(200, 50000, d, 7, clustered=True)
    vals = torch.randn(keys.shape[0], d, generator=torch.Generator().manual_seed(8))
    labs = torch.nn.functional.one_hot(torch.arange(keys.shape[0]) % 3, 3).float()
    base = R.ToyGraphBase(None, 3, d, 3, device=DEV, mode=L.SIM_TF32)
    base.retrieve_num = k
    base.add_entries(keys.to(DEV), vals.to(DEV), labs.to(DEV))
    s, i = base.topk(q.to(DEV), k)
    S64 = O.cosine_similarity_f64(q, keys)
    _, ref_i = O.topk(torch.from_numpy(S64), k)
    rows = [r for r in range(q.shape[0]) if r != 1]
    recall = O.recall_at_k(i.cpu().numpy()[rows], ref_i.numpy()[rows])
    print(f"tf32 recall@{k} on clustered keys: {recall:.4f}")
    assert recall >= 0.9                              # dense clusters: near-ties inside the tf32 error band may swap
    emb, lab = base.retrieve(q.to(DEV), None, False)
    assert torch.equal(emb.cpu(), vals[i.cpu()]) and torch.equal(lab.cpu(), labs[i.cpu()])


def test_tf32_unsupported_shapes_raise():
    lib = L.load()
    assert lib.rag_sim_mode_supported(L.SIM_TF32, 200, 10) == 0
    assert lib.rag_sim_mode_supported(L.SIM_TF32, 128, 26) == 0
    assert lib.rag_tf32_shadow_dpad(129) == 0 and lib.rag_tf32_shadow_dpad(33) == 64
    q = torch.randn(4, 200, device=DEV); keys = torch.randn(50, 200, device=DEV)
    with pytest.raises(RuntimeError):
        ops.rows_to_tf32(keys)
    with pytest.raises(RuntimeError):
        ops.cosine_topk(q, keys, 3, ops.row_inv_norm(keys), torch.zeros(50, 256, device=DEV), L.SIM_TF32)
