"""ncu target: the two-pass mode at the reference's Cora shape (2 708 queries x 10 832 keys x 256, k = 4), three exact-mode calls.
    ncu --set full --clock-control none -k regex:"cosine_topk_ts_kernel|refine2_kernel|sample_threshold" --launch-skip 8 --launch-count 4 \
        -o gpurun_out/r2_twopass_cfg1 python tools/r2_ncu_probe_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ragraph_b200 import _lib as L, ops

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
Q, N, d, k = 2708, 10832, 256, 4
keys = torch.randn(N, d, device=dev, generator=g); q = torch.randn(Q, d, device=dev, generator=g)
inv = ops.row_inv_norm(keys)
err = torch.zeros(1, device=dev)
sh, _ = ops.rows_to_shadow16(keys, L.FMT_F16, True, err_max=err)
torch.cuda.synchronize()
for _ in range(3):
    s, i = ops.cosine_topk(q, keys, k, inv, sh, L.SIM_F16_REFINE, 0, 0, err)
torch.cuda.synchronize()
print("ok", float(s[0, 0]))
