"""Round-2 probe: fp16 vs bf16 filter, kp 16 vs 32, clustered / duplicate libraries, pass-2 rows (one GPU).
    python tools/r2_probe.py [N] [d]        -> JSON lines on stdout"""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ragraph_b200 import _lib as L, ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
Q, k = 4096, 10
dev = "cuda"


def make(kind, seed=1234, chunk=2_000_000):
    keys = torch.empty(N, d, device=dev)
    g = torch.Generator(device=dev).manual_seed(seed)
    if kind == "gauss":
        for a in range(0, N, chunk):
            b = min(N, a + chunk); keys[a:b] = torch.randn(b - a, d, generator=g, device=dev)
        q = torch.randn(Q, d, generator=g, device=dev)
    else:
        cent = torch.randn(1024, d, generator=g, device=dev)
        for a in range(0, N, chunk):
            b = min(N, a + chunk)
            keys[a:b] = cent[torch.randint(0, 1024, (b - a,), generator=g, device=dev)] + 0.1 * torch.randn(b - a, d, generator=g, device=dev)
        q = cent[torch.randint(0, 1024, (Q,), generator=g, device=dev)] + 0.1 * torch.randn(Q, d, generator=g, device=dev)
        if kind == "dup":                       # clustered + 5 % exact duplicate rows
            src = torch.randint(0, N, (N // 20,), generator=g, device=dev)
            dst = torch.randint(0, N, (N // 20,), generator=g, device=dev)
            keys[dst] = keys[src]
    for a in range(0, N, chunk):
        b = min(N, a + chunk); keys[a:b] = torch.nn.functional.normalize(keys[a:b], dim=-1)
    return q, keys


def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


for kind in ("gauss", "clustered", "dup"):
    q, keys = make(kind)
    inv = ops.row_inv_norm(keys)
    s0 = i0 = None
    for mode, fmt in ((L.SIM_F16_REFINE, L.FMT_F16), (L.SIM_BF16_REFINE, L.FMT_BF16)):
        err = torch.zeros(1, device=dev)
        sh, _ = ops.rows_to_shadow16(keys, fmt, True, err_max=err)
        for kp in (16, 32) if d <= 128 else (16,):
            L.tc_set_option("kp", kp)
            if mode == L.SIM_BF16_REFINE and (kp == 32 or kind != "gauss") and N > 20_000_000:
                continue
            s, i, st = ops.cosine_topk_with_stats(q, keys, k, inv, sh, mode, shadow_err=err)
            med, mn = timeit(lambda: ops.cosine_topk(q, keys, k, inv, sh, mode, 0, 0, err))
            rec = {"kind": kind, "N": N, "d": d, "mode": mode, "kp": kp, "ms": round(med, 3), "ms_min": round(mn, 3),
                   "tflops": round(2 * Q * N * d / med / 1e9, 1), "pass2_rows": int(st[0]), "fp32_rows": int(st[1]), "kerr": float(err)}
            if s0 is None:
                # reference answer on 128 sampled rows: fp32 CUDA-core kernel
                rows = torch.arange(0, Q, Q // 128, device=dev)[:128]
                s0, i0 = ops.cosine_topk(q[rows].contiguous(), keys, k, inv)
                rec["ref"] = "fp32 kernel on 128 rows"
            rec["set_mismatch_rows"] = int((i[rows] != i0).any(dim=1).sum())
            rec["max_score_diff"] = float((s[rows] - s0).abs().max())
            print(json.dumps(rec), flush=True)
        L.tc_set_option("kp", -1)
        del sh
    del keys, q, inv
    torch.cuda.empty_cache()
