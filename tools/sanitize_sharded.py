"""torchrun worker for compute-sanitizer on the sharded finish kernel (world 2, tiny shapes):
    compute-sanitizer --tool memcheck --target-processes all python -m torch.distributed.run --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29533 tools/sanitize_sharded.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import ragraph_b200 as R
from ragraph_b200 import _lib as L

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator().manual_seed(1)
N, d, C = 6000, 64, 3
keys = torch.nn.functional.normalize(torch.randn(N, d, generator=g), dim=-1)
vals = torch.randn(N, d, generator=g)
labs = torch.nn.functional.one_hot(torch.randint(0, C, (N,), generator=g), C).float()
lo, hi = R.shard_bounds(N, world, rank)
st = R.ToyGraphBase(None, C, d, 3, device=dev, capacity=hi - lo, mode=L.SIM_FP32)
st.add_entries(keys[lo:hi].to(dev), vals[lo:hi].to(dev), labs[lo:hi].to(dev))
sr = R.ShardedRetriever(st, N)
for step in range(4):
    q = torch.randn(96, d, generator=g).to(dev)
    emb, lab, s, i = sr.retrieve(q, 5)
    assert sr.last_path == "p2p", sr.last_path
    assert torch.equal(emb.cpu(), vals[i.cpu()])
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    print("sanitize_sharded done: path", sr.last_path, flush=True)
dist.destroy_process_group()
