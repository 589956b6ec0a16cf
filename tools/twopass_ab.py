"""A/B of the two-pass mode for short key streams (maxima pass + collect pass instead of list warm-up) on one GPU:
    python tools/twopass_ab.py [Q N d k]...     default: the reference-sized shapes
Event-timed exact mode 5 with the mode off / on (automatic kernel choice), results compared."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from ragraph_b200 import _lib as L, ops

dev = torch.device("cuda", 0)
if os.environ.get("TWOPASS_MAX_TILES"):
    L.tc_set_option("twopass_max_tiles", int(os.environ["TWOPASS_MAX_TILES"]))
args = [int(x) for x in sys.argv[1:]]
cfgs = list(zip(args[0::4], args[1::4], args[2::4], args[3::4])) or [
    (2708, 10832, 256, 4), (2708, 43328, 256, 4), (4096, 240000, 64, 10), (300, 20000, 128, 10), (4096, 1000000, 128, 10), (512, 400000, 128, 10)]
for Q, N, d, k in cfgs:
    g = torch.Generator(device=dev).manual_seed(Q + N)
    keys = torch.randn(N, d, device=dev, generator=g)
    q = torch.randn(Q, d, device=dev, generator=g)
    inv = ops.row_inv_norm(keys)
    err = torch.zeros(1, device=dev)
    sh = ops.rows_to_shadow16(keys, L.FMT_F16, True, err_max=err)[0]
    res, times = {}, {0: [], 1: []}
    for tp in (0, 1):
        L.tc_set_option("twopass", tp)
        s, i, st = ops.cosine_topk_with_stats(q, keys, k, inv, sh, L.SIM_F16_REFINE, shadow_err=err)
        res[tp] = (s.clone(), i.clone(), st.tolist())
    for rep in range(5):
        for tp in (0, 1):
            L.tc_set_option("twopass", tp)
            times[tp].append(B.timeit_events(lambda: ops.cosine_topk(q, keys, k, inv, sh, L.SIM_F16_REFINE, 0, 0, err), 20, 5))
    L.tc_set_option("twopass", -1)
    med = {tp: sorted(v)[len(v) // 2] for tp, v in times.items()}
    print(json.dumps({"Q": Q, "N": N, "d": d, "k": k, "one_pass_us": round(med[0] * 1e3, 1), "two_pass_us": round(med[1] * 1e3, 1),
                      "speedup": round(med[0] / med[1], 3), "stats_one": res[0][2][:2], "stats_two": res[1][2][:2],
                      "max_score_diff": float((res[0][0] - res[1][0]).abs().max()),
                      "rows_idx_identical": float((res[0][1] == res[1][1]).all(dim=1).float().mean())}), flush=True)
    del keys, q, sh
    torch.cuda.empty_cache()
