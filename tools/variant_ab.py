"""A/B of the tensor-core filter kernel variants on ONE GPU (event-timed medians):

    python tools/variant_ab.py [N d]...     default: 12.5 M x 128 and 10 M x 256; Q = 4096, k = 10 (AB_Q / AB_K override)

variant ts = query tile stationary in tensor memory, ss = query tile in shared memory; for each: exact mode 3, raw
mode 2, MMA only (RAG_TC_DEBUG=1: epilogue skips TMEM reads) and MMA + TMEM loads without the filter (=2).
"""
import json, os, sys

os.environ["RAG_DIAG"] = "1"      # this tool uses the library's diagnostic switches (RAG_TC_DEBUG / trace / ...)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ragraph_b200 import ops, _lib as L

dev, Q, k = "cuda", int(os.environ.get("AB_Q", 4096)), int(os.environ.get("AB_K", 10))


def med(fn, iters=12, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return ms[len(ms) // 2]


def run(N, d):
    torch.manual_seed(0)
    keys = torch.empty(N, d, device=dev)
    for a in range(0, N, 4_000_000):
        b = min(N, a + 4_000_000)
        keys[a:b] = torch.nn.functional.normalize(torch.randn(b - a, d, device=dev), dim=-1)
    q = torch.randn(Q, d, device=dev)
    inv = ops.row_inv_norm(keys); shadow = ops.rows_to_bf16(keys, True)
    flop = 2.0 * Q * N * d
    res = {}
    for variant in VARIANTS:
        L.tc_set_option("variant", {"ss": 1, "ts": 2}[variant])
        out = {"N": N, "d": d, "Q": Q, "k": k, "variant": variant}
        for name, mode, dbg in (("exact_mode3", 3, None), ("raw_mode2", 2, None), ("mma_only", 2, "1"), ("mma_tmemld", 2, "2"), ("filter_no_hits", 2, "4")):
            if dbg is None:
                os.environ.pop("RAG_TC_DEBUG", None)
            else:
                os.environ["RAG_TC_DEBUG"] = dbg
            ms = med(lambda: ops.cosine_topk(q, keys, k, key_inv_norm=inv, keys_bf16=shadow, mode=mode))
            out[name + "_ms"] = round(ms, 4); out[name + "_tflops"] = round(flop / ms / 1e9, 1)
        os.environ.pop("RAG_TC_DEBUG", None)
        res[variant] = ops.cosine_topk(q, keys, k, key_inv_norm=inv, keys_bf16=shadow, mode=3)
        print(json.dumps(out), flush=True)
    if len(res) == 2:
        same = all(bool(torch.equal(res["ss"][j], res["ts"][j])) for j in (0, 1))
        print(json.dumps({"N": N, "d": d, "ss_equals_ts_mode3": same}), flush=True)
    L.tc_set_option("variant", -1)


VARIANTS = ("ss", "ts")

if __name__ == "__main__":
    if "--ts" in sys.argv:
        VARIANTS = ("ts",); sys.argv.remove("--ts")
    args = [int(x) for x in sys.argv[1:]]
    cfgs = list(zip(args[0::2], args[1::2])) or [(12_500_000, 128), (10_000_000, 256)]
    for N, d in cfgs:
        run(N, d)
