#!/usr/bin/env python
"""bench.py -- RAGraph hot path on B200: retrieval queries/sec (top-10, 100 M keys) + SpMM GB/s.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels behind the C ABI)
    python bench.py --impl reference --gpus N ...            # the reference's CPU torch path (oracle port)

A "step" is one 4 096-query batch through retrieve (fused similarity + top-10 over the key-row-sharded
library, NCCL all-gather merge when N > 1, gather of the winners' values and labels).  The library
(100 M x 128 fp32 keys + values + labels, synthetic, seeded) is resident in HBM like the reference's
``resource_keys.cuda()``; `value` times steps with the query batch already on the device, `e2e` times the
same call with HOST query/result buffers (pinned), copies inside the timed region.  The second half of the
metric, CSR SpMM on the ogbn-products-shaped synthetic graph, is reported in the same JSON line under
"spmm" (rank 0; replicas only).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

Q_BATCH, TOPK, DIM, N_KEYS, N_CLASS = 4096, 10, 128, 100_000_000, 3
SPMM_N, SPMM_NNZ, SPMM_F = 2_449_029, 61_859_140, 256
METRIC = "retrieval queries/sec (top-10, 100M keys) + SpMM GB/s"


def _ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from a committed `ncu --set full` capture of exactly this
    kernel and size (profiles/ncu_traffic.json, written from the .ncu-rep by tools/ncu_summary.py); None if not captured."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    return json.load(open(p)).get(key)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "bf16": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """SM clock + throttle reasons during the timed region (pynvml; one sample per 100 ms)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                    nv.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = get(self.h)
                self.reasons |= {k for k, b in names.items() if mask & b}
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------------- synthetic data
def make_library_shard(lo, hi, dim, n_class, device, seed=1234, chunk=4_000_000):
    """rows [lo, hi) of the synthetic library; every row depends only on (seed, chunk id) so any sharding
    of the same N yields the same global library."""
    import ragraph_b200 as R
    store = R.ToyGraphBase(None, n_class, dim, 3, device=device, capacity=hi - lo)
    store.retrieve_num = TOPK
    for cid in range(lo // chunk, (hi + chunk - 1) // chunk):
        a = cid * chunk
        g = torch.Generator(device=device).manual_seed(seed + cid)
        keys = torch.randn(chunk, dim, generator=g, device=device)
        keys = torch.nn.functional.normalize(keys, dim=-1)      # keys are normalised at insert (ToyGraphBase.py:109)
        vals = torch.randn(chunk, dim, generator=g, device=device)
        labs = torch.nn.functional.one_hot(torch.randint(0, n_class, (chunk,), generator=g, device=device), n_class).float()
        s, e = max(lo, a) - a, min(hi, a + chunk) - a
        store.add_entries(keys[s:e], vals[s:e], labs[s:e])
        del keys, vals, labs
    return store


def make_queries(Q, dim, device, seed=4321):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(Q, dim, generator=g).pin_memory() if device != "cpu" else torch.randn(Q, dim, generator=g)


def make_products_graph(device, n=SPMM_N, nnz=SPMM_NNZ, seed=7):
    """ogbn-products-shaped synthetic CSR: Chung-Lu endpoints with w_i ~ (i + 1.35)^-0.5 (mean degree 25.3, hubs
    up to ~17 k), node ids scattered by a random permutation, sym-norm weights 1/sqrt(d_i d_j)."""
    g = torch.Generator(device=device).manual_seed(seed)
    i0 = 1.35
    lo, hi = i0 ** 0.5, (n + i0) ** 0.5

    def endpoints():
        u = torch.rand(nnz, generator=g, device=device, dtype=torch.float64)
        return (((u * (hi - lo) + lo) ** 2 - i0).clamp_(0, n - 1)).long()

    perm = torch.randperm(n, generator=g, device=device)
    dst = perm[endpoints()]
    src = perm[endpoints()]
    deg_in = torch.bincount(dst, minlength=n)
    deg_out = torch.bincount(src, minlength=n)
    dst, order = torch.sort(dst)
    src = src[order]
    del order
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    torch.cumsum(deg_in, 0, out=rowptr[1:])
    val = (deg_in[dst].clamp(min=1).float() * deg_out[src].clamp(min=1).float()).rsqrt_()
    return rowptr, src.to(torch.int32), val, int(deg_in.max())


# ----------------------------------------------------------------------------------------- reference / CPU baseline
def cpu_retrieve_rate(n_sample, steps, warmup):
    """The reference's torch CPU path (oracle port of ToyGraphBase.retrieve) on a bounded library sample;
    returns seconds per 4 096-query batch at n_sample keys."""
    from oracle import ragraph_oracle as O
    g = torch.Generator().manual_seed(1234)
    keys = torch.nn.functional.normalize(torch.randn(n_sample, DIM, generator=g), dim=-1)
    vals = torch.randn(n_sample, DIM, generator=g)
    labs = torch.nn.functional.one_hot(torch.randint(0, N_CLASS, (n_sample,), generator=g), N_CLASS).float()
    q = make_queries(Q_BATCH, DIM, "cpu")
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        O.retrieve(q, keys, vals, labs, TOPK)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = args.ref_sample
    t = cpu_retrieve_rate(n_sample, args.steps, args.warmup)
    qps = Q_BATCH / (t * (N_KEYS / n_sample))
    cores = torch.get_num_threads()
    sample = (f"oracle port of ToyGraphBase.retrieve, torch CPU fp32, Q={Q_BATCH} x N={n_sample} keys d={DIM} per step; "
              f"q/s extrapolated linearly in N to {N_KEYS} keys")
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3 * (N_KEYS / n_sample),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"top-{TOPK} cosine retrieve, {N_KEYS} keys d={DIM}, {Q_BATCH}-query batches (CPU sample {n_sample} keys)"},
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


class StdoutToStderr:
    """File-descriptor level: until restore(), everything written to stdout -- C libraries included -- lands on stderr.
    NCCL prints its version banner straight to stdout when NCCL_DEBUG=VERSION (NCCL_DEBUG_FILE is only honoured from
    WARN upwards), which would put a second line in front of the ONE JSON line the driver reads."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def restore(self):
        if self.saved is not None:
            sys.stdout.flush()
            os.dup2(self.saved, 1)
            os.close(self.saved)
            self.saved = None


# ----------------------------------------------------------------------------------------- our arm
def main():
    global N_KEYS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-keys", type=int, default=N_KEYS)
    ap.add_argument("--mode", type=int, default=-1, help="-1 auto, 0 fp32 CUDA cores, 2 bf16 raw, 3 bf16 filter + fp32 refine")
    ap.add_argument("--ref-sample", type=int, default=250_000)
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    ap.add_argument("--no-spmm", action="store_true")
    ap.add_argument("--nccl-exchange", action="store_true", help="multi-GPU: all-gather formulation instead of the peer-memory kernel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        N_KEYS = args.n_keys
        return run_reference(args)

    # rank 0 prints ONE JSON line on stdout: NCCL's own log (version banner at VERSION/WARN level) goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    guard = StdoutToStderr()
    import torch.distributed as dist
    import ragraph_b200 as R
    from ragraph_b200 import _lib as L, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N_KEYS = args.n_keys
    peaks = _peaks()

    lo, hi = R.shard_bounds(N_KEYS, world, rank)
    store = make_library_shard(lo, hi, DIM, N_CLASS, dev)
    if args.mode >= 0:
        store.mode = args.mode
    if args.nccl_exchange:
        os.environ["RAG_P2P"] = "0"
    sr = R.ShardedRetriever(store, N_KEYS)
    mode = store._pick_mode(Q_BATCH, TOPK)
    q_host = make_queries(Q_BATCH, DIM, dev)
    q_dev = q_host.to(dev)
    out_host = {"emb": torch.empty((Q_BATCH, TOPK, DIM), dtype=torch.float32).pin_memory(),
                "lab": torch.empty((Q_BATCH, TOPK, N_CLASS), dtype=torch.float32).pin_memory(),
                "idx": torch.empty((Q_BATCH, TOPK), dtype=torch.int64).pin_memory(),
                "sc": torch.empty((Q_BATCH, TOPK), dtype=torch.float32).pin_memory()}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]

    def step(q, it=None):
        # local fused top-k (event-timed for the roofline) + the sharded finish; N=1: plain gathers
        emb, lab, scores, idx = sr.retrieve(q, TOPK, copy=False, events=k_ev[it] if it is not None else None)
        return emb, lab, scores, idx

    # ---- device-resident timing ("value") --------------------------------------------------
    for _ in range(args.warmup):
        step(q_dev)
    barrier()
    launches0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for it in range(args.steps):
            res = step(q_dev, it)
        e1.record()
        barrier()
    launches = L.launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    kern_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in k_ev) / args.steps)
    qps = Q_BATCH / (ms_step * 1e-3)

    # ---- end-to-end with host buffers ("e2e") ----------------------------------------------
    def e2e_step():
        q = q_host.to(dev, non_blocking=True)
        emb, lab, scores, idx = step(q)
        out_host["emb"].copy_(emb, non_blocking=True); out_host["lab"].copy_(lab, non_blocking=True)
        out_host["idx"].copy_(idx, non_blocking=True); out_host["sc"].copy_(scores, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    h2d = q_host.numel() * 4
    d2h = sum(t.numel() * t.element_size() for t in out_host.values())

    # ---- sanity: the timed path's answer is right (exact fp32 scores of returned ids, sorted, in range) --
    emb, lab, scores, idx = res
    chk = min(64, Q_BATCH)
    assert bool((scores[:, :-1] >= scores[:, 1:]).all()) and bool(((idx >= 0) & (idx < N_KEYS)).all())
    mine = (idx[:chk] >= lo) & (idx[:chk] < hi)
    loc = (idx[:chk] - lo).clamp(0, hi - lo - 1)
    kk = store.resource_keys[loc.reshape(-1)].double().reshape(chk, TOPK, DIM)
    ex = (torch.nn.functional.normalize(q_dev[:chk].double(), dim=-1)[:, None] * torch.nn.functional.normalize(kk, dim=-1)).sum(-1)
    tol = 1e-5 if mode not in (L.SIM_BF16, L.SIM_F16) else 1e-2
    assert float(((scores[:chk].double() - ex).abs() * mine).max()) < tol, "returned scores disagree with exact re-score"
    assert bool(torch.equal(emb[:chk][mine], store.resource_values[loc[mine]])), "gather not bit exact"

    flops = 2.0 * Q_BATCH * (hi - lo) * DIM
    tf_ach = flops / (kern_ms * 1e-3) / 1e12
    roof = {"bound": "tensor", "achieved": tf_ach, "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": tf_ach / peaks["bf16"],
            "traffic": _ncu_traffic(f"cosine_topk_ts_kernel:N={hi - lo}:d={DIM}:Q={Q_BATCH}") if mode in (2, 3, 4, 5) else None,
            "kernel": {0: "cosine_topk_f32_kernel (CUDA-core fp32)",
                       2: "cosine_topk_ts_kernel (tcgen05 bf16, query tile stationary in TMEM) + threshold pre-pass",
                       3: "cosine_topk_ts_kernel (tcgen05 bf16, query tile stationary in TMEM) + threshold pre-pass + fp32 refine",
                       4: "cosine_topk_ts_kernel (tcgen05 fp16, query tile stationary in TMEM) + threshold pre-pass",
                       5: "cosine_topk_ts_kernel (tcgen05 fp16, query tile stationary in TMEM) + threshold pre-pass + fp32 refine"
                       }.get(mode, str(mode)),
            "kernel_ms": kern_ms, "peak_source": peaks["src"] + " bf16 burst (cuBLAS 8192^3)",
            "peak_sustained": peaks["bf16_sustained"], "frac_of_sustained": tf_ach / peaks["bf16_sustained"],
            "algorithmic": "2*Q*N_local*d flop per launch"}

    line = {"metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": {0: "f32", 2: "bf16", 3: "bf16 filter + f32 refine (exact)", 4: "f16",
                                         5: "f16 filter + f32 refine (exact)"}.get(mode, "f32"),
            "data": "synthetic",
            "config": {"workload": f"top-{TOPK} cosine retrieve + value/label gather, {N_KEYS} keys d={DIM} sharded by key rows over "
                                   f"{world} GPU(s), {Q_BATCH}-query batches", "mode": mode,
                       "l2": "inputs larger than L2 (key shard streamed every step)", "parallelism": f"key-row shard x{world}",
                       "exchange": {"p2p": "one kernel over NVLink peer memory (push candidates, merge, owners store rows to peers)",
                                    "nccl": "NCCL all-gather + merge + owner gather + all-gather", "single": "none"}[sr.last_path]},
            "roofline": roof,
            "e2e": {"value": Q_BATCH / (e2e_ms * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms},
            "gpu_launches": int(launches), "clocks": clk.summary()}

    # ---- SpMM half of the metric (rank 0, replicas only) ------------------------------------
    if not args.no_spmm and rank == 0:
        del res, emb, lab
        try:
            line["spmm"] = bench_spmm(dev, args, peaks)
        except torch.cuda.OutOfMemoryError as e:            # library shard + graph do not both fit
            line["spmm"] = {"skipped": f"OOM next to the library shard: {str(e)[:80]}"}
    if not args.no_cpu_baseline and rank == 0 and world == 1:
        n_s = args.cpu_sample
        t = cpu_retrieve_rate(n_s, 2, 1)
        line["cpu_baseline"] = {"value": Q_BATCH / (t * (N_KEYS / n_s)), "unit": "queries/s", "cores": torch.get_num_threads(),
                                "kind": "port", "sample": f"oracle port of ToyGraphBase.retrieve (torch CPU fp32), Q={Q_BATCH} x N={n_s} "
                                f"keys d={DIM}, mean of 2 runs; q/s extrapolated linearly in N to {N_KEYS}"}
    if world > 1:
        dist.barrier()
    guard.restore()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        guard = StdoutToStderr()            # teardown chatter stays off stdout too
        dist.destroy_process_group()


def bench_spmm(dev, args, peaks):
    from ragraph_b200 import _lib as L, ops
    rowptr, col, val, max_deg = make_products_graph(dev)
    x = torch.randn(SPMM_N, SPMM_F, device=dev)
    steps = max(args.steps, 5)
    for _ in range(3):
        y = ops.csr_spmm(rowptr, col, val, x)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    l0 = L.launch_count()
    for a, b in ev:
        a.record(); y = ops.csr_spmm(rowptr, col, val, x); b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    # property check at full size: A.1 == row sums of val
    ones = ops.csr_spmm(rowptr, col, val, torch.ones(SPMM_N, 16, device=dev))[:, 0].double()
    rs = torch.zeros(SPMM_N, device=dev, dtype=torch.float64).index_add_(
        0, torch.repeat_interleave(torch.arange(SPMM_N, device=dev), rowptr[1:] - rowptr[:-1]), val.double())
    assert float((ones - rs).abs().max() / rs.abs().max()) < 1e-5, "SpMM row-sum property failed"
    alg = SPMM_NNZ * 8 + (SPMM_N + 1) * 8 + SPMM_NNZ * SPMM_F * 4 + SPMM_N * SPMM_F * 4
    gbs = alg / (ms * 1e-3) / 1e9
    return {"value": gbs, "unit": "GB/s (algorithmic, edge-gather model)", "ms": ms, "edges_per_s": SPMM_NNZ / (ms * 1e-3),
            "config": {"workload": f"CSR SpMM ogbn-products-shaped synthetic: n={SPMM_N} nnz={SPMM_NNZ} F={SPMM_F} fp32, "
                                   f"Chung-Lu power law, max degree {max_deg}", "l2": "X (2.5 GB) larger than L2"},
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                         "traffic": _ncu_traffic(f"csr_spmm_kernel:n={SPMM_N}:nnz={SPMM_NNZ}:F={SPMM_F}"),
                         "kernel": "csr_spmm_kernel<32,2>", "algorithmic_bytes": alg,
                         "compulsory_bytes": SPMM_NNZ * 8 + (SPMM_N + 1) * 8 + 2 * SPMM_N * SPMM_F * 4,
                         "peak_source": peaks["src"] + " copy bandwidth"},
            "gpu_launches": int(L.launch_count() - l0)}


if __name__ == "__main__":
    main()
