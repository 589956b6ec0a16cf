"""A/B at full shard size on one GPU: filter format (fp16 / bf16) x list length (kp) x library kind.
    python tools/r2_ab.py N d kinds(comma) configs(comma of mode:kp)"""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from ragraph_b200 import _lib as L, ops

N, d = int(sys.argv[1]), int(sys.argv[2])
kinds = sys.argv[3].split(",")
cfgs = [tuple(int(x) for x in c.split(":")) for c in sys.argv[4].split(",")]
dev = torch.device("cuda", 0)
Q, k = 4096, 10
store = B.make_library_shard(0, N, d, 3, dev, kinds[0])
for kind in kinds:
    if kind != kinds[0]:
        B.fill_library_shard(store, 0, N, d, 3, dev, kind)
    q = B.make_queries(Q, d, dev, kind=kind).to(dev)
    keys, inv = store.resource_keys, store.key_inv_norm
    rows = torch.arange(0, Q, 32, device=dev)
    s0, i0 = ops.cosine_topk(q[rows].contiguous(), keys, k, inv)
    shadows = {}
    for mode, kp in cfgs:
        fmt = L.FMT_F16 if mode in (4, 5) else L.FMT_BF16
        if fmt not in shadows:
            shadows.clear(); torch.cuda.empty_cache()
            err = torch.zeros(1, device=dev)
            shadows[fmt] = (ops.rows_to_shadow16(keys, fmt, True, err_max=err)[0], err)
        sh, err = shadows[fmt]
        L.tc_set_option("kp", kp)
        s, i, st = ops.cosine_topk_with_stats(q, keys, k, inv, sh, mode, shadow_err=err)
        ts = []
        for rep in range(3):
            ts.append(B.timeit_events(lambda: ops.cosine_topk(q, keys, k, inv, sh, mode, 0, 0, err), 5, 2))
        ms = sorted(ts)[1]
        differ = (i[rows] != i0).any(dim=1)
        print(json.dumps({"kind": kind, "N": N, "d": d, "mode": mode, "kp": kp, "ms": round(ms, 3), "ms_all": [round(t, 3) for t in ts],
                          "tflops": round(2 * Q * N * d / ms / 1e9, 1), "pass2_rows": int(st[0]), "fp32_rows": int(st[1]),
                          "rows_differ_vs_fp32": int(differ.sum()), "max_score_diff": float((s[rows] - s0).abs().max())}), flush=True)
    L.tc_set_option("kp", -1)
