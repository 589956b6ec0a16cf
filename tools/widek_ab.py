"""k in (26, 128] on the tensor cores (more key splits, not longer lists) vs the fp32 CUDA-core kernel and stock torch:
    python tools/widek_ab.py [Q N d k]...    default: the edge variant's vanilla phase (32768 x 240000 x 64, k = 50) and 4096 x 2 M x 128, k = 100"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import bench as B
from ragraph_b200 import _lib as L, ops

dev = torch.device("cuda", 0)
args = [int(x) for x in sys.argv[1:]]
cfgs = list(zip(args[0::4], args[1::4], args[2::4], args[3::4])) or [(32768, 240000, 64, 50), (4096, 2_000_000, 128, 100)]
for Q, N, d, k in cfgs:
    g = torch.Generator(device=dev).manual_seed(Q + N)
    keys = torch.randn(N, d, device=dev, generator=g)
    q = torch.randn(Q, d, device=dev, generator=g)
    inv = ops.row_inv_norm(keys)
    err = torch.zeros(1, device=dev)
    sh = ops.rows_to_shadow16(keys, L.FMT_F16, True, err_max=err)[0]
    s5, i5, st = ops.cosine_topk_with_stats(q, keys, k, inv, sh, L.SIM_F16_REFINE, shadow_err=err)
    s0, i0 = ops.cosine_topk(q, keys, k, inv)
    chunk = max(1, min(Q, (6 << 30) // (4 * N)))

    def stock():
        kn = F.normalize(keys, dim=-1)
        outs = []
        for a in range(0, Q, chunk):
            sim = F.normalize(q[a:a + chunk], dim=-1) @ kn.T
            outs.append(torch.topk(sim, k, largest=True, sorted=True))
        return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])

    ss, si = stock()
    rows_diff = int((i5 != i0).any(dim=1).sum())
    ms_tc = B.timeit_events(lambda: ops.cosine_topk(q, keys, k, inv, sh, L.SIM_F16_REFINE, 0, 0, err), 5, 2)
    ms_f32 = B.timeit_events(lambda: ops.cosine_topk(q, keys, k, inv), 3, 1)
    ms_stock = B.timeit_events(stock, 3, 1)
    print(json.dumps({"Q": Q, "N": N, "d": d, "k": k, "tc_f16_refine_ms": round(ms_tc, 3), "fp32_kernel_ms": round(ms_f32, 3),
                      "stock_torch_ms": round(ms_stock, 3), "tc_tflops": round(2 * Q * N * d / ms_tc / 1e9, 1),
                      "pass2_rows": int(st[0]), "fp32_rows": int(st[1]), "rows_idx_differ_vs_fp32_kernel": rows_diff,
                      "max_score_diff_vs_fp32_kernel": float((s5 - s0).abs().max()),
                      "max_score_diff_vs_stock": float((s5 - ss).abs().max()),
                      "rows_idx_differ_vs_stock": int((i5 != si).any(dim=1).sum())}), flush=True)
    del keys, q, sh, ss, si
    torch.cuda.empty_cache()
