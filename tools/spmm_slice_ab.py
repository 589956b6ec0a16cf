"""A/B of the SpMM's feature-sliced schedule at cfg5 (ogbn-products-shaped graph, F = 256) on one GPU:
    python tools/spmm_slice_ab.py [slice:hints ...]        default: 0:0 128:1 64:1 32:1 32:0
slice = floats per column slice (0 = one launch over whole rows), hints = L2 eviction-priority hints on / off.
Interleaved event timing; every configuration's result is compared with the one-launch result."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from ragraph_b200 import _lib as L, ops

dev = torch.device("cuda", 0)
cfgs = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(0, 0), (128, 1), (64, 1), (32, 1), (32, 0)]
rowptr, col, val, max_deg = B.make_products_graph(dev)
x = torch.randn(B.SPMM_N, B.SPMM_F, device=dev)
alg = B.SPMM_NNZ * 8 + (B.SPMM_N + 1) * 8 + B.SPMM_NNZ * B.SPMM_F * 4 + B.SPMM_N * B.SPMM_F * 4


def setcfg(c):
    L.spmm_set_option("slice", c[0]); L.spmm_set_option("l2_hints", c[1])


setcfg((0, 0))
y0 = ops.csr_spmm(rowptr, col, val, x)
y0e = ops.csr_spmm(rowptr, col, val, x, epilogue=L.EPI_ROWNORM | L.EPI_RELU)
outs = []
for c in cfgs:
    setcfg(c)
    y = ops.csr_spmm(rowptr, col, val, x)
    ye = ops.csr_spmm(rowptr, col, val, x, epilogue=L.EPI_ROWNORM | L.EPI_RELU)
    fin = torch.isfinite(y0e)
    outs.append({"slice": c[0], "l2_hints": c[1], "bit_identical": bool(torch.equal(y, y0)),
                 "max_rel_diff": float((y - y0).abs().max() / y0.abs().max()),
                 "epi_max_rel_diff": float((ye[fin] - y0e[fin]).abs().max() / y0e[fin].abs().max()),
                 "epi_nan_pattern_same": bool(torch.equal(torch.isnan(ye), torch.isnan(y0e)))})
    del y, ye
times = [[] for _ in cfgs]
for rep in range(4):
    for i, c in enumerate(cfgs):
        setcfg(c)
        times[i].append(B.timeit_events(lambda: ops.csr_spmm(rowptr, col, val, x), 5, 2))
for o, ts in zip(outs, times):
    ms = sorted(ts)[len(ts) // 2]
    o.update({"ms": round(ms, 3), "ms_all": [round(t, 3) for t in ts], "algorithmic_GBps": round(alg / ms / 1e6, 1),
              "edges_per_s": round(B.SPMM_NNZ / ms * 1e3)})
    print(json.dumps(o), flush=True)
L.spmm_set_option("slice", -1); L.spmm_set_option("l2_hints", -1)
