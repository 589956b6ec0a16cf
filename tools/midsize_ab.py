"""Mid-size retrieval shapes (the reference's real ones): tensor-core kernel variants vs the fp32 path, event-timed.

    python tools/midsize_ab.py
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ragraph_b200 import ops, _lib as L

dev = "cuda"
SHAPES = [(4096, 240_000, 64, 10, "edge variant batch (modules/RAGraph.py:298-324)"),
          (4096, 1_000_000, 128, 10, "1 M keys"),
          (4096, 4_000_000, 128, 10, "4 M keys"),
          (2708, 10_832, 256, 4, "cfg1 Cora-shaped node"),
          (512, 12_000, 256, 4, "node batch of 16 TU graphs"),
          (1, 480, 256, 3, "cfg2 graph variant")]


def med(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return ms[len(ms) // 2]


for Q, N, d, k, what in SHAPES:
    torch.manual_seed(0)
    keys = torch.nn.functional.normalize(torch.randn(N, d, device=dev), dim=-1); q = torch.randn(Q, d, device=dev)
    inv = ops.row_inv_norm(keys); shadow = ops.rows_to_bf16(keys, True)
    out = {"Q": Q, "N": N, "d": d, "k": k, "what": what}
    out["fp32_ms"] = round(med(lambda: ops.cosine_topk(q, keys, k, key_inv_norm=inv)), 4)
    for name, env in (("ss", {"variant": 1}), ("ts", {"variant": 2}), ("ts_pre64", {"variant": 2, "prepass_min_tiles": 64})):
        for kk in ("variant", "prepass_min_tiles"):
            L.tc_set_option(kk, -1)
        for kk, vv in env.items():
            L.tc_set_option(kk, vv)
        try:
            out[name + "_ms"] = round(med(lambda: ops.cosine_topk(q, keys, k, key_inv_norm=inv, keys_bf16=shadow, mode=3)), 4)
        except Exception as e:
            out[name + "_ms"] = str(e)[:60]
    print(json.dumps(out), flush=True)
