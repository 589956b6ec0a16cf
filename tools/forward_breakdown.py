"""Kernel-by-kernel time of one RAGraph.forward at BASELINE config 1 / 2 (ncu launch list target):
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fwd.csv python tools/forward_breakdown.py node
then  python tools/forward_breakdown.py --read gpurun_out/fwd.csv  prints the last forward's kernels."""
import csv, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 2 and sys.argv[1] == "--read":
    rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 5 and r[0].isdigit()]
    names = [(r[4], float(r[-1])) for r in rows]             # Kernel Name, value (ns or us as ncu prints; unit column before)
    unit = rows[0][-2] if rows else "?"
    # the script runs 4 identical forwards after warm-up markers; print the last quarter
    n = len(names) // 4
    tot = 0.0
    for nm, v in names[-n:]:
        print(f"{v:10.2f} {unit}  {nm[:110]}")
        tot += v
    print(f"{tot:10.2f} {unit}  total of {n} kernels")
    sys.exit(0)

import torch
import torch.nn.functional as F
import ragraph_b200 as R

variant = sys.argv[1] if len(sys.argv) > 1 else "node"
dev = torch.device("cuda", 0)
n, f, d, C, N, k = (2708, 1433, 256, 7, 10832, 4) if variant == "node" else (40, 3, 256, 2, 480, 3)
g = torch.Generator(device=dev).manual_seed(3)
a = (torch.rand(n, n, generator=g, device=dev) < 3.9 / n).float()
a = torch.triu(a, 1); a = a + a.t() + torch.eye(n, device=dev)
dinv = a.sum(1).pow(-0.5); adj = (dinv[:, None] * a * dinv[None, :]).contiguous()
x = torch.randn(n, f, generator=g, device=dev)


class Enc(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.gcn = R.GCN(f, d, "prelu")

    def inference(self, features, adj_):
        return self.gcn([features, adj_])


base = R.ToyGraphBase(None, C, d, 3, device=dev, capacity=N, variant=variant)
base.retrieve_num = k
base.add_entries(F.normalize(torch.randn(N, d, generator=g, device=dev), dim=-1), torch.randn(N, d, generator=g, device=dev),
                 F.one_hot(torch.randint(0, C, (N,), generator=g, device=dev), C).float())
model = R.RAGraph(Enc(), base, f, C, d, variant=variant).to(dev).eval()
with torch.no_grad():
    model(x, adj)                       # warm-up: CSR conversion, shadows (not part of the 4 measured forwards)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(4):
        model(x, adj)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok")
