"""ncu target, round 2: N d fmt(bf16|fp16) [kind] -- three exact-mode calls of the fused similarity + top-k at one shard shape.
    ncu --set full --clock-control none --import-source on -k regex:cosine_topk_ts_kernel --launch-skip 3 --launch-count 2 \
        -o gpurun_out/r2_ts_12m5 python tools/r2_ncu_probe.py 12500000 128 bf16"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from ragraph_b200 import _lib as L, ops

N, d, fmt = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
kind = sys.argv[4] if len(sys.argv) > 4 else "gauss"
dev = torch.device("cuda", 0)
store = B.make_library_shard(0, N, d, 3, dev, kind)
q = B.make_queries(4096, d, dev, kind=kind).to(dev)
mode, f = (L.SIM_BF16_REFINE, L.FMT_BF16) if fmt == "bf16" else (L.SIM_F16_REFINE, L.FMT_F16)
err = torch.zeros(1, device=dev)
sh, _ = ops.rows_to_shadow16(store.resource_keys, f, True, err_max=err)
inv = store.key_inv_norm
torch.cuda.synchronize()
for _ in range(3):
    s, i = ops.cosine_topk(q, store.resource_keys, 10, inv, sh, mode, 0, 0, err)
torch.cuda.synchronize()
print("ok", float(s[0, 0]))
