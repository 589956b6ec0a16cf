"""Small-shape pass through every kernel of the tensor-core retrieval path (SS filter, TS filter + pre-pass, refine, second
pass + refine2, fp32 fallback, gathers) for compute-sanitizer:
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_probe.py
Shapes are tiny so that the 10-100x sanitizer slowdown stays in seconds; results are checked against the fp32 kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ragraph_b200 import _lib as L, ops

dev = "cuda"
g = torch.Generator().manual_seed(0)
d, k = 128, 10
cent = torch.randn(6, d, generator=g)
N, Q = 40000, 300
keys = (cent[torch.randint(0, 6, (N,), generator=g)] + 0.1 * torch.randn(N, d, generator=g)).to(dev)
q = (cent[torch.randint(0, 6, (Q,), generator=g)] + 0.1 * torch.randn(Q, d, generator=g)).to(dev)
keys[100:140] = keys[7]                                   # duplicates: ties + spill
inv = ops.row_inv_norm(keys)
s0, i0 = ops.cosine_topk(q, keys, k, inv)
L.tc_set_option("prepass_min_tiles", 64)
for mode, fmt in ((L.SIM_F16_REFINE, L.FMT_F16), (L.SIM_BF16_REFINE, L.FMT_BF16)):
    err = torch.zeros(1, device=dev)
    sh, _ = ops.rows_to_shadow16(keys, fmt, True, err_max=err)
    for variant in (1, 2):
        L.tc_set_option("variant", variant)
        s, i, st = ops.cosine_topk_with_stats(q, keys, k, inv, sh, mode, shadow_err=err)
        torch.cuda.synchronize()
        assert float((s - s0).abs().max()) < 2e-6, (mode, variant)
        print(f"mode {mode} variant {variant}: ok, pass2 rows {int(st[0])}, fp32 rows {int(st[1])}", flush=True)
# cross-split threshold sharing (sweeping CTAs next to the workers) and k > 26 (more key splits): 10 query tiles x 14 splits
L.tc_set_option("variant", 2)
Q2, N2, d2 = 2400, 14 * 64 * 128, 64
keys2 = torch.randn(N2, d2, generator=g).to(dev); q2 = torch.randn(Q2, d2, generator=g).to(dev)
inv2 = ops.row_inv_norm(keys2)
err2 = torch.zeros(1, device=dev)
sh2, _ = ops.rows_to_shadow16(keys2, L.FMT_F16, True, err_max=err2)
rows = torch.arange(0, Q2, 8, device=dev)
for kk in (10, 50):
    if kk == 50:          # 32-entry lists: a shorter stream (racecheck slows the kernel ~100x; its barrier waits trap after 10 s)
        q2, keys2, inv2, sh2 = q2[:520].contiguous(), keys2[:40000].contiguous(), inv2[:40000].contiguous(), sh2[:40000].contiguous()
        rows = torch.arange(0, 520, 8, device=dev)
    s, i, st = ops.cosine_topk_with_stats(q2, keys2, kk, inv2, sh2, L.SIM_F16_REFINE, shadow_err=err2)
    sr, ir = ops.cosine_topk(q2[rows].contiguous(), keys2, kk, inv2)
    torch.cuda.synchronize()
    assert float((s[rows] - sr).abs().max()) < 2e-6, kk
    print(f"sweep / wide-k probe k={kk}: ok, worker CTAs seen by the sweep {int(st[3])}, hits queued {int(st[4])}", flush=True)
L.tc_set_option("variant", -1)
# two-pass mode (automatic kernel choice on short streams): Gaussian keys and the clustered library (overflow -> retry list)
for nm, (qq, kk_, iv) in {"gauss": (q2[:300].contiguous(), keys2, inv2), "clustered": (q, keys, inv)}.items():
    e3 = torch.zeros(1, device=dev)
    s3h, _ = ops.rows_to_shadow16(kk_, L.FMT_F16, True, err_max=e3)
    s, i, st = ops.cosine_topk_with_stats(qq, kk_, 10, iv, s3h, L.SIM_F16_REFINE, shadow_err=e3)
    sr, ir = ops.cosine_topk(qq, kk_, 10, iv)
    torch.cuda.synchronize()
    assert float((s - sr).abs().max()) < 2e-6, nm
    print(f"two-pass probe {nm}: ok, retried rows {int(st[0])}, fp32 rows {int(st[1])}", flush=True)
vals = torch.randn(N, d, device=dev)
out = ops.gather_rows(vals, i0)
assert torch.equal(out, vals[i0])
red = ops.gather_reduce(vals, i0, 0)
torch.cuda.synchronize()
print("sanitize_probe done", flush=True)
