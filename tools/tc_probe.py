"""Time the fused retrieval op alone (GPU): python tools/tc_probe.py [N] [d] [Q] [mode] [iters]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ragraph_b200 import ops, _lib as L

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
Q = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 3
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 10
k = 10
dev = "cuda"
torch.manual_seed(0)
keys = torch.empty(N, d, device=dev)
for a in range(0, N, 4_000_000):
    b = min(N, a + 4_000_000)
    keys[a:b] = torch.nn.functional.normalize(torch.randn(b - a, d, device=dev), dim=-1)
q = torch.randn(Q, d, device=dev)
inv = ops.row_inv_norm(keys)
shadow = ops.rows_to_bf16(keys, True) if mode != 0 else None
for _ in range(3):
    s, i = ops.cosine_topk(q, keys, k, key_inv_norm=inv, keys_bf16=shadow, mode=mode)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
for a, b in ev:
    a.record(); s, i = ops.cosine_topk(q, keys, k, key_inv_norm=inv, keys_bf16=shadow, mode=mode); b.record()
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in ev)
med = ms[len(ms) // 2]
print(f"RAG_TC_DEBUG={os.environ.get('RAG_TC_DEBUG','0')} N={N} d={d} Q={Q} mode={mode}: median {med:.3f} ms  min {ms[0]:.3f}  "
      f"{2*Q*N*d/med/1e9:.1f} TFLOP/s  {Q/med*1e3:.0f} q/s")
if os.environ.get("RAG_TC_DEBUG") == "3":
    import ctypes, numpy as np
    lib = L.load()
    buf = (ctypes.c_ulonglong * (2048 + 512))()
    lib.rag_tc_trace_read.argtypes = [ctypes.c_void_p]
    lib.rag_tc_trace_read(buf)
    tr = np.array(buf[:2048], dtype=np.int64).reshape(512, 4)
    tr2 = np.frombuffer(np.array(buf[2048:], dtype=np.uint64).tobytes(), dtype=np.uint32).reshape(512, 2)
    t0 = tr[0, 0]
    print("tile: mma_start mma_issued | epi_start epi_done   (cycles since first)   d(mma_start) d(epi_done)")
    for t in list(range(0, 12)) + list(range(400, 412)):
        r = tr[t] - t0
        dm = tr[t, 0] - tr[t - 1, 0] if t else 0
        de = tr[t, 3] - tr[t - 1, 3] if t else 0
        print(f"{t:4d}: {r[0]:8d} {r[1]:8d} | {r[2]:8d} {r[3]:8d}    {dm:6d} {de:6d}   epi_len={tr[t,3]-tr[t,2]}  full_lat={tr[t,2]-tr[t,1]}  drains={tr2[t,0]}  tmem_hold={tr2[t,1]}")
    print("drains per tile (100..500):", float(tr2[100:500, 0].mean()), " mean tmem_hold:", float(tr2[100:500, 1].mean()),
          " mean epi_len for drains==0:", float(np.mean([(tr[t,3]-tr[t,2]) for t in range(100,500) if tr2[t,0]==0] or [0])),
          " ==1:", float(np.mean([(tr[t,3]-tr[t,2]) for t in range(100,500) if tr2[t,0]==1] or [0])),
          " ==2:", float(np.mean([(tr[t,3]-tr[t,2]) for t in range(100,500) if tr2[t,0]==2] or [0])))
    print("mean period (tiles 100..500):", (tr[500, 0] - tr[100, 0]) / 400.0, " mean epi_len:", float(np.mean(tr[100:500, 3] - tr[100:500, 2])))
