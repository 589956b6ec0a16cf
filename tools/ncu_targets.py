"""Small driver for ncu captures: runs each hot kernel a few times at a representative size.
    ncu --set full -k regex:... python tools/ncu_targets.py [tc|spmm|gather]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ragraph_b200 import ops, _lib as L
which = sys.argv[1:] or ["tc", "spmm", "gather"]
dev = "cuda"
torch.manual_seed(0)
if "tc" in which:
    N, d, Q, k = 20_000_000, 128, 4096, 10
    keys = torch.empty(N, d, device=dev)
    for a in range(0, N, 4_000_000):
        keys[a:a + 4_000_000] = torch.nn.functional.normalize(torch.randn(min(4_000_000, N - a), d, device=dev), dim=-1)
    q = torch.randn(Q, d, device=dev)
    inv = ops.row_inv_norm(keys); shadow = ops.rows_to_bf16(keys, True)
    for _ in range(4):
        ops.cosine_topk(q, keys, k, key_inv_norm=inv, keys_bf16=shadow, mode=3)
    torch.cuda.synchronize()
    del keys, shadow
if "spmm" in which:
    from bench import make_products_graph, SPMM_N, SPMM_F
    rowptr, col, val, _ = make_products_graph(dev)
    x = torch.randn(SPMM_N, SPMM_F, device=dev)
    for _ in range(4):
        ops.csr_spmm(rowptr, col, val, x)
    torch.cuda.synchronize()
    del rowptr, col, val, x
if "gather" in which:
    N, d, Q, k = 2_449_029, 256, 2_400_000, 10
    table = torch.randn(N, d, device=dev); idx = torch.randint(0, N, (Q, k), device=dev)
    for _ in range(4):
        ops.gather_rows(table, idx)
    for _ in range(4):
        ops.gather_reduce(table, idx, L.REDUCE_MEAN)
    torch.cuda.synchronize()
