"""Time the fused retrieval op alone (GPU): python tools/tc_probe.py [N] [d] [Q] [mode] [iters]"""
import os, sys, time

os.environ["RAG_DIAG"] = "1"      # this tool uses the library's diagnostic switches (RAG_TC_DEBUG / trace / ...)
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
print(f"variant={os.environ.get('RAG_TC_VARIANT','ts')} RAG_TC_DEBUG={os.environ.get('RAG_TC_DEBUG','0')} N={N} d={d} Q={Q} mode={mode}: median {med:.3f} ms  min {ms[0]:.3f}  "
      f"{2*Q*N*d/med/1e9:.1f} TFLOP/s  {Q/med*1e3:.0f} q/s")
if os.environ.get("RAG_TC_DEBUG") == "3" or os.environ.get("RAG_TC_TRACE") == "1":
    import ctypes, numpy as np
    lib = L.load()
    buf = (ctypes.c_ulonglong * (2048 + 512 + 2048))()
    lib.rag_tc_trace_read.argtypes = [ctypes.c_void_p]
    lib.rag_tc_trace_read(buf)
    tr = np.array(buf[:2048], dtype=np.int64).reshape(512, 4)
    tr2 = np.frombuffer(np.array(buf[2048:2560], dtype=np.uint64).tobytes(), dtype=np.uint32).reshape(512, 2)
    tr3 = np.frombuffer(np.array(buf[2560:], dtype=np.uint64).tobytes(), dtype=np.uint32).reshape(256, 16).astype(np.int64)
    t0 = tr[0, 0]
    print("tile: mma_start mma_issued | epi_start epi_done   (cycles since first)   d(mma_start) d(epi_done)")
    for t in list(range(0, 12)) + list(range(400, 412)):
        r = tr[t] - t0
        dm = tr[t, 0] - tr[t - 1, 0] if t else 0
        de = tr[t, 3] - tr[t - 1, 3] if t else 0
        print(f"{t:4d}: {r[0]:8d} {r[1]:8d} | {r[2]:8d} {r[3]:8d}    {dm:6d} {de:6d}   epi_len={tr[t,3]-tr[t,2]}  full_lat={tr[t,2]-tr[t,1]}  drains={tr2[t,0]}  tmem_hold={tr2[t,1]}")
    hits = tr2[100:500, 1]
    el = (tr[100:500, 3] - tr[100:500, 2])
    per = np.diff(tr[100:501, 0])
    print(f"window t0={os.environ.get('RAG_TC_TRACE_T0','0')}: mean period {per.mean():.1f} (median {np.median(per):.0f}, p90 {np.percentile(per,90):.0f})  "
          f"issue {np.mean(tr[100:500,1]-tr[100:500,0]):.0f}")
    for h in (0, 1, 2, 3):
        sel = hits == h if h < 3 else hits >= 3
        if sel.any():
            print(f"  warp-0 lanes on the slow path = {h}{'+' if h == 3 else ''}: {sel.mean()*100:.1f}% of tiles, epi_len mean {el[sel].mean():.0f} median {np.median(el[sel]):.0f}, "
                  f"period mean {per[sel].mean():.0f}")
    names = ["top", "W_e0", "fulls", "mma0", "-", "mma1", "commit0", "W_e1", "mma_rb1", "commit1", "empties"]
    seg = (tr3[100:250, 1:11] - tr3[100:250, 0:10]) & 0xffffffff
    nxt = (tr3[101:251, 0] - tr3[100:250, 10]) & 0xffffffff
    print("MMA-thread micro-trace, mean cycles per segment (tiles 100..250): " +
          "  ".join(f"{n}={seg[:, i].mean():.0f}" for i, n in enumerate(names[1:])) + f"  loop={nxt.mean():.0f}  total={(seg.sum(1) + nxt).mean():.0f}")
