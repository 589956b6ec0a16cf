"""A/B of the cross-split threshold sharing (sweeping CTAs on the idle SMs, TcArgs::pool) on one GPU:
    python tools/gshare_ab.py N d kinds(comma) modes(comma)      e.g. 12500000 128 gaussian,clustered 3,5
For each (kind, mode): event-timed median with the sweep off / on, rows in the second pass, and the results of both
runs compared bit for bit (both are exact) plus sampled rows against the fp32 CUDA-core kernel."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from ragraph_b200 import _lib as L, ops

N, d = int(sys.argv[1]), int(sys.argv[2])
kinds = sys.argv[3].split(",")
modes = [int(x) for x in sys.argv[4].split(",")]
# sweep configurations "ctas[:prepass_div[:dbg]]", e.g. 4,2,4:128,4:64:1
def _cfg(x):
    p = [int(v) for v in x.split(":")]
    return (p[0], p[1] if len(p) > 1 else 64, p[2] if len(p) > 2 else 0)
ctas = [_cfg(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [(4, 64, 0)]
dev = torch.device("cuda", 0)


def setcfg(g, c):
    L.tc_set_option("gshare", g); L.tc_set_option("gshare_ctas", c[0]); L.tc_set_option("prepass_div", c[1]); L.tc_set_option("gshare_dbg", c[2])


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
    for mode in modes:
        fmt = L.FMT_F16 if mode in (4, 5) else L.FMT_BF16
        if fmt not in shadows:
            shadows.clear(); torch.cuda.empty_cache()
            err = torch.zeros(1, device=dev)
            shadows[fmt] = (ops.rows_to_shadow16(keys, fmt, True, err_max=err)[0], err)
        sh, err = shadows[fmt]
        ref = None
        cfgs = [(0, (4, 64, 0))] + [(1, c) for c in ctas]
        outs = []
        for g, nc in cfgs:
            setcfg(g, nc)
            s, i, st = ops.cosine_topk_with_stats(q, keys, k, inv, sh, mode, shadow_err=err)
            st5 = st.tolist()
            out = {"kind": kind, "N": N, "d": d, "mode": mode, "gshare": g, "sweep_ctas": nc[0] if g else 0, "prepass_div": nc[1], "dbg": nc[2],
                   "pass2_rows": int(st5[0]), "fp32_rows": int(st5[1]), "hits_queued": int(st5[4]) & 0xffffffff,
                   "rows_differ_vs_fp32": int((i[rows] != i0).any(dim=1).sum()), "max_score_diff_vs_fp32": float((s[rows] - s0).abs().max())}
            if ref is None:
                ref = (s, i)
            else:
                out["identical_to_sweep_off"] = bool(torch.equal(i, ref[1]) and torch.equal(s, ref[0]))
            outs.append(out)
        # interleaved timing: the power cap makes back-to-back blocks of one configuration drift by several per cent
        times = [[] for _ in cfgs]
        for rep in range(5):
            for ci, (g, nc) in enumerate(cfgs):
                setcfg(g, nc)
                times[ci].append(B.timeit_events(lambda: ops.cosine_topk(q, keys, k, inv, sh, mode, 0, 0, err), 6, 2))
        for out, ts in zip(outs, times):
            ms = sorted(ts)[len(ts) // 2]
            out.update({"ms": round(ms, 3), "ms_all": [round(t, 3) for t in ts], "tflops": round(2 * Q * N * d / ms / 1e9, 1)})
            print(json.dumps(out), flush=True)
    L.tc_set_option("gshare", -1); L.tc_set_option("gshare_ctas", -1); L.tc_set_option("prepass_div", -1); L.tc_set_option("gshare_dbg", 0)
