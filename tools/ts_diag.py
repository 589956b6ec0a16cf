"""Diagnostic (GPU): raw bf16 scores of the two tensor-core kernel variants against the exact fp64 cosine.

    python tools/ts_diag.py            prints max |score - exact| and recall@10 for ss / ts / ts with swapped halves
"""
import os, sys

os.environ["RAG_DIAG"] = "1"      # this tool uses the library's diagnostic switches (RAG_TC_DEBUG / trace / ...)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ragraph_b200 import ops, _lib as L

dev = "cuda"
for (Q, N, d, k) in [(256, 4096, 128, 10), (300, 9000, 64, 10), (300, 9000, 256, 10), (300, 9000, 160, 10)]:
    torch.manual_seed(0)
    q = torch.randn(Q, d, device=dev); keys = torch.randn(N, d, device=dev)
    inv = ops.row_inv_norm(keys); shadow = ops.rows_to_bf16(keys, True)
    S = torch.nn.functional.normalize(q.double(), dim=-1) @ torch.nn.functional.normalize(keys.double(), dim=-1).T
    ref_s, ref_i = S.topk(k, dim=1)
    for variant, swap in (("ss", "0"), ("ts", "0"), ("ts", "1")):
        L.tc_set_option("variant", {"ss": 1, "ts": 2}[variant]); os.environ["RAG_TS_SWAP"] = swap
        try:
            s, i = ops.cosine_topk(q, keys, k, key_inv_norm=inv, keys_bf16=shadow, mode=L.SIM_BF16)
            torch.cuda.synchronize()
        except Exception as e:
            print(f"Q={Q} N={N} d={d} variant={variant} swap={swap}: ERROR {str(e)[:200]}", flush=True)
            continue
        ok = i >= 0
        ex = S.gather(1, i.clamp(min=0))
        err = float(((s.double() - ex).abs() * ok).max())
        rec = float((i[:, :, None] == ref_i[:, None, :]).any(-1).float().mean())
        print(f"Q={Q} N={N} d={d} variant={variant} swap={swap}: max|s-exact|={err:.2e} recall@{k}={rec:.4f} invalid={int((~ok).sum())}", flush=True)
os.environ.pop("RAG_TS_SWAP", None)
