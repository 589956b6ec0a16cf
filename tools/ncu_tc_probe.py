"""One-kernel driver for ncu source-level captures of the tensor-core filter: python tools/ncu_tc_probe.py [N] [d] [launches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ragraph_b200 import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
torch.manual_seed(0)
keys = torch.empty(N, d, device="cuda")
for a in range(0, N, 4_000_000):
    b = min(N, a + 4_000_000)
    keys[a:b] = torch.nn.functional.normalize(torch.randn(b - a, d, device="cuda"), dim=-1)
q = torch.randn(4096, d, device="cuda")
inv = ops.row_inv_norm(keys); shadow = ops.rows_to_bf16(keys, True)
for _ in range(n):
    ops.cosine_topk(q, keys, 10, key_inv_norm=inv, keys_bf16=shadow, mode=3)
torch.cuda.synchronize()
