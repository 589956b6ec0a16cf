"""Where does one retrieve step go?  Event-timed phases of the sharded step on ONE GPU at a given shard size
(default 12.5 M keys = the 8-GPU shard of the 100 M-key library).

    python tools/step_breakdown.py [N ...]        one JSON line per N
"""
import json
import os
import sys

os.environ["RAG_DIAG"] = "1"      # this tool uses the library's diagnostic switches (RAG_TC_DEBUG / trace / ...)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ragraph_b200 import ops, _lib as L

dev = "cuda"
d, Q, k, C = 128, 4096, 10, 3


def med(fn, iters=20, warmup=4):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return ms[len(ms) // 2]


def run(N):
    torch.manual_seed(0)
    keys = torch.empty(N, d, device=dev)
    for a in range(0, N, 4_000_000):
        b = min(N, a + 4_000_000)
        keys[a:b] = torch.nn.functional.normalize(torch.randn(b - a, d, device=dev), dim=-1)
    vals = torch.randn(N, d, device=dev) if N <= 30_000_000 else keys
    labs = torch.zeros(N, C, device=dev)
    q = torch.randn(Q, d, device=dev)
    inv = ops.row_inv_norm(keys)
    shadow = ops.rows_to_bf16(keys, True)
    out = {"N": N, "Q": Q, "d": d, "k": k}
    flop = 2.0 * Q * N * d

    def topk(mode):
        return ops.cosine_topk(q, keys, k, key_inv_norm=inv, keys_bf16=shadow, mode=mode)

    for name, mode, dbg in (("exact_mode3", 3, None), ("raw_mode2", 2, None), ("mma_only_dbg1", 2, "1"),
                            ("mma_tmemld_dbg2", 2, "2")):
        if dbg is None:
            os.environ.pop("RAG_TC_DEBUG", None)
        else:
            os.environ["RAG_TC_DEBUG"] = dbg
        ms = med(lambda: topk(mode))
        out[name + "_ms"] = round(ms, 4)
        out[name + "_tflops"] = round(flop / ms / 1e9, 1)
    os.environ.pop("RAG_TC_DEBUG", None)
    s, i = topk(3)
    out["gather_values_ms"] = round(med(lambda: ops.gather_rows(vals, i)), 4)
    out["gather_labels_ms"] = round(med(lambda: ops.gather_rows(labs, i)), 4)
    out["q_prologue_ms"] = round(med(lambda: (ops.rows_to_bf16(q, True), ops.row_inv_norm(q))), 4)
    out["empty_launch_pair_ms"] = round(med(lambda: (ops.row_inv_norm(q[:8]), ops.row_inv_norm(q[:8]))), 4)
    # host-side cost of one op call (python + ctypes + tensor-map encode), GPU idle
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        topk(3)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    out["host_issue_ms_per_call"] = round((t1 - t0) / 20 * 1e3, 4)
    print(json.dumps(out), flush=True)
    del keys, shadow, inv


if __name__ == "__main__":
    for n in [int(x) for x in sys.argv[1:]] or [12_500_000]:
        run(n)
