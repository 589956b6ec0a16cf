#!/usr/bin/env python
"""bench.py -- RAGraph hot path on B200: retrieval queries/sec (top-10, 100 M keys) + SpMM GB/s.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels behind the C ABI)
    python bench.py --impl reference --gpus N ...            # the reference's CPU torch path (oracle port)

A "step" is one 4 096-query batch through retrieve (fused similarity + top-10 over the key-row-sharded library, peer-memory
merge when N > 1, gather of the winners' values and labels).  The library (100 M x 128 fp32 keys + values + labels,
synthetic, seeded) is resident in HBM like the reference's ``resource_keys.cuda()``; `value` times steps with the query
batch already on the device, `e2e` times the same call with HOST query / result buffers (pinned), copies inside the timed
region (result copy of step i overlapped with step i+1; every rank copies back its own slice of the query rows).

Before anything is timed the answer of the timed path is checked at full size (`parity`): sampled query rows against a full
fp32 scan by the CUDA-core kernel AND by stock torch (matmul + topk, the reference's own call sequence), tie-aware through
fp64 scores of the union of the id sets.  The same JSON line carries: `variants` (clustered / 5 %-duplicate libraries of
SURVEY 8d, with the rows that needed the second tensor-core pass or the fp32 kernel), `gpu_stock_baseline` (the reference's
torch calls on this B200), `spmm` (the second half of the metric, ogbn-products-shaped CSR), `gather`, `cfg3` (10 M x 256),
`small` (the reference's real shapes), `cpu_baseline`.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

Q_BATCH, TOPK, DIM, N_KEYS, N_CLASS = 4096, 10, 128, 100_000_000, 3
N_CENTROIDS, CLUSTER_SIGMA = 1024, 0.1
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
    """SM clock + throttle reasons during the timed region (pynvml; one sample per 10 ms -- the timed region of an 8-GPU run is
    ~90 ms; the sample taken as the region opens is dropped when later ones exist: it still sees the idle clock)."""

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
            self._stop.wait(0.01)

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
        s = sorted(self.samples[1:] if len(self.samples) > 3 else self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------------- synthetic data
def _centroids(dim, device, seed):
    g = torch.Generator(device="cpu").manual_seed(seed + 99991)
    return torch.randn(N_CENTROIDS, dim, generator=g).to(device)


def fill_library_shard(store, lo, hi, dim, n_class, device, kind="gauss", seed=1234, chunk=4_000_000):
    """Rows [lo, hi) of the synthetic library into `store` (emptied first).  Every row depends only on (seed, chunk id), so
    any sharding of the same N yields the same global library.
      gauss     keys ~ N(0, I) (SURVEY 8d cfg3/cfg4 base case)
      clustered keys = centroid + 0.1 N(0, I), 1 024 Gaussian centroids (the sigma of Augmentation.augment_features,
                RAGraph_node/ragraph_utils/Augmentation.py:8-20): ~N/1024 near neighbours per query
      dup5      clustered + 5 % of the rows are exact copies of another row (key, value and label), as the reference's
                multinomial(replacement=True) draw produces (ToyGraphBase.py:98)
    Keys are L2-normalised at insert like the reference (ToyGraphBase.py:109)."""
    store.clear()
    cent = _centroids(dim, device, seed) if kind != "gauss" else None
    for cid in range(lo // chunk, (hi + chunk - 1) // chunk):
        a = cid * chunk
        g = torch.Generator(device=device).manual_seed(seed + cid)
        keys = torch.randn(chunk, dim, generator=g, device=device)
        vals = torch.randn(chunk, dim, generator=g, device=device)
        labs = F.one_hot(torch.randint(0, n_class, (chunk,), generator=g, device=device), n_class).float()
        if cent is not None:
            keys = keys.mul_(CLUSTER_SIGMA).add_(cent[torch.randint(0, N_CENTROIDS, (chunk,), generator=g, device=device)])
        keys = F.normalize(keys, dim=-1)
        if kind == "dup5":
            dst = torch.randint(0, chunk, (chunk // 20,), generator=g, device=device)
            src = torch.randint(0, chunk, (chunk // 20,), generator=g, device=device)
            keys[dst] = keys[src]; vals[dst] = vals[src]; labs[dst] = labs[src]
        s, e = max(lo, a) - a, min(hi, a + chunk) - a
        store.add_entries(keys[s:e], vals[s:e], labs[s:e])
        del keys, vals, labs
    return store


def make_library_shard(lo, hi, dim, n_class, device, kind="gauss", seed=1234):
    import ragraph_b200 as R
    store = R.ToyGraphBase(None, n_class, dim, 3, device=device, capacity=hi - lo)
    store.retrieve_num = TOPK
    return fill_library_shard(store, lo, hi, dim, n_class, device, kind, seed)


def make_queries(Q, dim, device, seed=4321, kind="gauss", lib_seed=1234):
    """host (pinned) query batch; for the clustered libraries the queries are perturbed cluster members too"""
    g = torch.Generator(device="cpu").manual_seed(seed)
    q = torch.randn(Q, dim, generator=g)
    if kind != "gauss":
        cent = _centroids(dim, "cpu", lib_seed)
        q = q * CLUSTER_SIGMA + cent[torch.randint(0, N_CENTROIDS, (Q,), generator=g)]
    return q.pin_memory() if str(device) != "cpu" else q


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


def timeit_events(fn, iters, warmup=3):
    """mean of per-call CUDA-event times on the current stream"""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / iters


# ----------------------------------------------------------------------------------------- reference / CPU baseline
def cpu_retrieve_time(n_sample, steps, warmup):
    """The reference's torch CPU path (oracle port of ToyGraphBase.retrieve) on a bounded library sample;
    returns seconds per 4 096-query batch at n_sample keys."""
    from oracle import ragraph_oracle as O
    g = torch.Generator().manual_seed(1234)
    keys = F.normalize(torch.randn(n_sample, DIM, generator=g), dim=-1)
    vals = torch.randn(n_sample, DIM, generator=g)
    labs = F.one_hot(torch.randint(0, N_CLASS, (n_sample,), generator=g), N_CLASS).float()
    q = make_queries(Q_BATCH, DIM, "cpu")
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        O.retrieve(q, keys, vals, labs, TOPK)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def cpu_fit(samples, steps, warmup):
    """time the port at >= 2 library sizes, fit t = a + b N (the reference's cost is one sgemm + one topk over [Q, N]: linear in
    N once N is out of cache), extrapolate to N_KEYS.  Returns (seconds per batch at N_KEYS, description, cores, points)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)              # torchrun exports OMP_NUM_THREADS=1: the baseline gets every host core anyway
    pts = []
    for i, n in enumerate(samples):
        pts.append((n, cpu_retrieve_time(n, steps if i == 0 else max(2, steps // 5), warmup if i == 0 else 1)))
    if len(pts) >= 2 and pts[-1][0] != pts[0][0]:
        (n0, t0), (n1, t1) = pts[0], pts[-1]
        b = (t1 - t0) / (n1 - n0)
        a = t0 - b * n0
        if b <= 0:
            a, b = 0.0, t1 / n1
    else:
        a, b = 0.0, pts[0][1] / pts[0][0]
    t_full = a + b * N_KEYS
    desc = ("oracle port of ToyGraphBase.retrieve (torch CPU fp32, %d host threads), Q=%d d=%d timed at N = %s keys (%s s per "
            "batch); fit t = %.3g + %.3g N, extrapolated to N = %d" %
            (torch.get_num_threads(), Q_BATCH, DIM, ", ".join(str(n) for n, _ in pts), ", ".join("%.3f" % t for _, t in pts),
             a, b, N_KEYS))
    return t_full, desc, torch.get_num_threads(), pts


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_full, desc, cores, pts = cpu_fit([args.ref_sample, 4 * args.ref_sample], args.steps, args.warmup)
    qps = Q_BATCH / t_full
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_full * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"top-{TOPK} cosine retrieve, {N_KEYS} keys d={DIM}, {Q_BATCH}-query batches "
                                   f"(CPU samples {[n for n, _ in pts]} keys, linear fit)"},
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": desc},
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


# ----------------------------------------------------------------------------------------- stock torch on the GPU
def stock_topk(q, keys, k, chunk, idx_offset=0, n_limit=None):
    """The reference's call sequence on stock torch CUDA (SimilarityFunctions.py:6-16 + ToyGraphBase.py:67): F.normalize both
    sides (the keys again on every call, as the reference does), fp32 matmul, torch.topk -- key-chunked because [Q, N] does
    not fit, chunk results merged with one more topk.  Returns (scores, global idx)."""
    N = keys.shape[0] if n_limit is None else min(n_limit, keys.shape[0])
    qn = F.normalize(q, p=2, dim=-1)
    best_s = best_i = None
    for a in range(0, N, chunk):
        b = min(N, a + chunk)
        s = torch.matmul(qn, F.normalize(keys[a:b], p=2, dim=-1).t())
        ts, ti = torch.topk(s, min(k, b - a), dim=1, largest=True, sorted=True)
        ti = ti + (a + idx_offset)
        if best_s is None:
            best_s, best_i = ts, ti
        else:
            cs, ci = torch.cat([best_s, ts], 1), torch.cat([best_i, ti], 1)
            best_s, sel = torch.topk(cs, k, dim=1, largest=True, sorted=True)
            best_i = torch.gather(ci, 1, sel)
        del s
    return best_s, best_i


# ----------------------------------------------------------------------------------------- parity at full size
class Dist:
    def __init__(self, dist, world, rank, dev):
        self.dist, self.world, self.rank, self.dev = dist, world, rank, dev

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max(self, x: float) -> float:
        if self.world == 1:
            return x
        t = torch.tensor([x], device=self.dev, dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t

    def merge_topk(self, s, i, k):
        """[rows, k] per-rank candidates with global ids -> global top-k, plain torch (checker code)"""
        if self.world == 1:
            return s, i
        all_s = torch.empty((self.world,) + tuple(s.shape), dtype=s.dtype, device=s.device)
        all_i = torch.empty((self.world,) + tuple(i.shape), dtype=i.dtype, device=i.device)
        self.dist.all_gather_into_tensor(all_s, s.contiguous())
        self.dist.all_gather_into_tensor(all_i, i.contiguous())
        cs = all_s.permute(1, 0, 2).reshape(s.shape[0], -1)
        ci = all_i.permute(1, 0, 2).reshape(s.shape[0], -1)
        # deterministic order (score desc, index asc) like the library's merge
        order = torch.argsort(ci, dim=1, stable=True)
        cs, ci = torch.gather(cs, 1, order), torch.gather(ci, 1, order)
        order = torch.argsort(cs, dim=1, descending=True, stable=True)[:, :k]
        return torch.gather(cs, 1, order), torch.gather(ci, 1, order)


def parity_probe(D: Dist, store, lo, hi, q_dev, ours_s, ours_i, n_rows, n_stock_rows, stock_chunk):
    """The timed path's answer on sampled query rows against (a) a full fp32 scan by the CUDA-core kernel (mode 0), (b) the
    same scan by stock torch.  Tie-aware: a row mismatches only if, by the fp64 scores of the UNION of the id sets (owners
    compute, summed over ranks), a returned id scores more than 1e-6 below the k-th best or a better id is missing."""
    from ragraph_b200 import _lib as L, ops
    Q = q_dev.shape[0]
    rows = torch.arange(0, Q, max(1, Q // n_rows), device=q_dev.device)[:n_rows]
    qs = q_dev[rows].contiguous()
    keys = store.resource_keys
    s_a, i_a = ops.cosine_topk(qs, keys, TOPK, store.key_inv_norm, None, L.SIM_FP32, 0, lo)
    s_a, i_a = D.merge_topk(s_a, i_a, TOPK)
    srow = rows[:n_stock_rows]
    s_b, i_b = stock_topk(q_dev[srow].contiguous(), keys, TOPK, stock_chunk, idx_offset=lo)
    s_b, i_b = D.merge_topk(s_b, i_b, TOPK)

    def f64_scores(qrows, ids):
        mine = (ids >= lo) & (ids < hi)
        loc = (ids - lo).clamp(0, hi - lo - 1)
        kk = F.normalize(keys[loc.reshape(-1)].double(), dim=-1).reshape(ids.shape + (keys.shape[1],))
        qq = F.normalize(q_dev[qrows].double(), dim=-1)
        s = (qq[:, None, :] * kk).sum(-1) * mine
        return D.sum_(s)

    def compare(qrows, got_i, ref_i, got_s=None):
        union = torch.cat([got_i, ref_i], 1)
        s64 = f64_scores(qrows, union)
        k = got_i.shape[1]
        got64, ref64 = s64[:, :k], s64[:, k:]
        # the k-th best score of the union is at least the larger of the two sets' minima: an id more than 1e-6 below it
        # is a wrong answer (in either set), everything else is a tie
        kth = torch.maximum(got64.min(dim=1).values, ref64.min(dim=1).values)
        bad = (got64 < kth[:, None] - 1e-6).any(dim=1) | (ref64 < kth[:, None] - 1e-6).any(dim=1)
        err = float((got_s.double() - got64).abs().max()) if got_s is not None else None
        return int(bad.sum()), err, int((got_i != ref_i).any(dim=1).sum())

    mism_a, err, differ_a = compare(rows, ours_i[rows], i_a, ours_s[rows])
    mism_b, _, differ_b = compare(srow, ours_i[srow], i_b)
    mism_ab, _, _ = compare(srow, i_a[:n_stock_rows], i_b)
    return {"rows": int(rows.numel()), "set_mismatch": mism_a + mism_b, "max_score_err": err,
            "reference": "full fp32 scan of the sampled rows by cosine_topk_f32_kernel (mode 0)",
            "rows_with_tie_swaps": differ_a,
            "stock_torch_rows": int(srow.numel()), "stock_torch_set_mismatch": mism_b, "stock_rows_with_tie_swaps": differ_b,
            "fp32_kernel_vs_stock_torch_mismatch": mism_ab,
            "criterion": "index sets identical except ties within 1e-6 (fp64 scores of the union); scores within 1e-5"}


# ----------------------------------------------------------------------------------------- our arm
def main():
    global N_KEYS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-keys", type=int, default=N_KEYS)
    ap.add_argument("--mode", type=int, default=-1, help="-1 auto, 0 fp32 CUDA cores, 2/4 bf16/fp16 raw, 3/5 bf16/fp16 filter + fp32 refine")
    ap.add_argument("--data", default="gauss", choices=["gauss", "clustered", "dup5"], help="library distribution of the headline line")
    ap.add_argument("--ref-sample", type=int, default=250_000)
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    ap.add_argument("--parity-rows", type=int, default=256)
    ap.add_argument("--no-spmm", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip variants / cfg3 / gather / stock baselines / small shapes")
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
    D = Dist(dist, world, rank, dev)
    N_KEYS = args.n_keys
    peaks = _peaks()
    t_wall0 = time.perf_counter()

    lo, hi = R.shard_bounds(N_KEYS, world, rank)
    store = make_library_shard(lo, hi, DIM, N_CLASS, dev, args.data)
    if args.mode >= 0:
        store.mode = args.mode
    if args.nccl_exchange:
        os.environ["RAG_P2P"] = "0"
    sr = R.ShardedRetriever(store, N_KEYS)
    mode = store._pick_mode(Q_BATCH, TOPK)
    exact = mode in (L.SIM_FP32, L.SIM_BF16_REFINE, L.SIM_F16_REFINE)
    stock_chunk = 262_144

    def run_config(kind, steps, warmup, parity_rows, stock_rows, with_e2e):
        """parity probe, device-timed steps, (optionally) end-to-end steps on the library currently in `store`"""
        q_host = make_queries(Q_BATCH, DIM, dev, kind=kind)
        q_dev = q_host.to(dev)
        out = {}
        for _ in range(warmup):
            sr.retrieve(q_dev, TOPK, copy=False)
            torch.cuda.synchronize()        # untimed: lets the store's filter-format policy see each call's counters and settle
        # ---- parity first: the answer of the path that is about to be timed -----------------------------
        store.collect_stats = True
        emb, lab, scores, idx = sr.retrieve(q_dev, TOPK, copy=True)
        store.collect_stats = False
        st = store.last_stats.clone().long() if store.last_stats is not None else torch.zeros(2, dtype=torch.long, device=dev)
        st = D.sum_(st).tolist()
        assert bool((scores[:, :-1] >= scores[:, 1:]).all()) and bool(((idx >= 0) & (idx < N_KEYS)).all())
        if exact:
            par = parity_probe(D, store, lo, hi, q_dev, scores, idx, parity_rows, stock_rows, stock_chunk)
            par["pass2_rows"], par["fallback_rows"] = int(st[0]), int(st[1])
            par["note"] = ("pass2_rows: query rows (summed over ranks) whose first-pass certificate failed and that took the "
                           "second tensor-core pass; fallback_rows: rows recomputed by the fp32 kernel")
            out["parity"] = par
            assert par["set_mismatch"] == 0, f"parity probe failed: {par}"
            assert par["max_score_err"] < 1e-5, par
        # gathers are bit exact (rows this rank owns)
        chk = min(64, Q_BATCH)
        mine = (idx[:chk] >= lo) & (idx[:chk] < hi)
        loc = (idx[:chk] - lo).clamp(0, hi - lo - 1)
        assert bool(torch.equal(emb[:chk][mine], store.resource_values[loc[mine]])), "gather not bit exact"
        assert bool(torch.equal(lab[:chk][mine], store.resource_labels[loc[mine]])), "label gather not bit exact"
        del emb, lab

        # ---- device-resident timing ("value") ----------------------------------------------------------
        k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for _ in range(warmup):
            sr.retrieve(q_dev, TOPK, copy=False)
        D.barrier()
        state0 = store.auto_state() if args.mode < 0 else None
        launches0 = L.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk:
            e0.record()
            for it in range(steps):
                sr.retrieve(q_dev, TOPK, copy=False, events=k_ev[it])
            e1.record()
            D.barrier()
        launches = L.launch_count() - launches0
        ms_step = D.max(e0.elapsed_time(e1)) / steps
        kern_ms = D.max(sum(a.elapsed_time(b) for a, b in k_ev) / steps)
        out.update(ms_per_step=ms_step, kernel_ms=kern_ms, qps=Q_BATCH / (ms_step * 1e-3), launches=int(launches),
                   clocks=clk.summary(), exchange=sr.last_path, filter=state0,
                   filter_stable=(state0 == store.auto_state()) if state0 is not None else True)
        if not with_e2e:
            return out

        # ---- end-to-end with host buffers ("e2e") --------------------------------------------------------
        # Every step: H2D copy of the query batch (pinned), retrieve, D2H copy of this rank's slice of the results into
        # pinned host buffers.  The result copy of step i runs on a second stream under step i+1 (the peer-memory result
        # block is double buffered by step parity, so step i+2 waits for copy i); the host reads result i after its copy
        # event -- all of it inside the timed region.
        r0, r1 = R.shard_bounds(Q_BATCH, world, rank)            # this rank answers query rows [r0, r1)
        nb = 2
        host = [{"emb": torch.empty((r1 - r0, TOPK, DIM), dtype=torch.float32).pin_memory(),
                 "lab": torch.empty((r1 - r0, TOPK, N_CLASS), dtype=torch.float32).pin_memory(),
                 "idx": torch.empty((r1 - r0, TOPK), dtype=torch.int64).pin_memory(),
                 "sc": torch.empty((r1 - r0, TOPK), dtype=torch.float32).pin_memory()} for _ in range(nb)]
        copy_stream = torch.cuda.Stream(device=dev)
        main_stream = torch.cuda.current_stream()
        done = [None] * nb
        keep = [None] * nb
        checksum = [0.0]

        def e2e_step(i):
            b = i % nb
            if done[b] is not None:
                done[b].synchronize()                            # host consumes result i-2 ...
                checksum[0] += float(host[b]["sc"][0, 0])        # ... (reads the pinned buffer)
                main_stream.wait_event(done[b])                  # ... and its device block may be overwritten from here on
            q = q_host.to(dev, non_blocking=True)
            # copy i-1 must be complete before finish(i) releases the peers into step i+1 (same result-block parity)
            emb, lab, sc, ix = sr.retrieve(q, TOPK, copy=False, wait_event=done[(i - 1) % nb])
            ready = torch.cuda.Event()
            ready.record(main_stream)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ready)
                host[b]["emb"].copy_(emb[r0:r1], non_blocking=True); host[b]["lab"].copy_(lab[r0:r1], non_blocking=True)
                host[b]["idx"].copy_(ix[r0:r1], non_blocking=True); host[b]["sc"].copy_(sc[r0:r1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            done[b] = ev
            keep[b] = (emb, lab, sc, ix, q)                      # keep the device tensors alive until their copy is done

        def e2e_drain():
            for b in range(nb):
                if done[b] is not None:
                    done[b].synchronize()
                    checksum[0] += float(host[b]["sc"][0, 0])
                    done[b] = None

        for i in range(3):
            e2e_step(i)
        e2e_drain()
        D.barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            e2e_step(i)
        e2e_drain()
        D.barrier()
        e2e_ms = D.max((time.perf_counter() - t0) * 1e3) / steps
        assert bool(torch.equal(host[(steps - 1) % nb]["idx"], idx[r0:r1].cpu())), "e2e result differs from the device-timed one"
        out["e2e"] = {"value": Q_BATCH / (e2e_ms * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": q_host.numel() * 4,
                      "d2h_bytes_per_step": sum(t.numel() * t.element_size() for t in host[0].values()),
                      "ms_per_step": e2e_ms,
                      "note": f"per rank: full query batch in, result rows [{r0}, {r1}) out; copy of step i overlaps step i+1"}
        return out

    main_run = run_config(args.data, args.steps, args.warmup, args.parity_rows, 32, True)
    ms_step, kern_ms, qps = main_run["ms_per_step"], main_run["kernel_ms"], main_run["qps"]

    flops = 2.0 * Q_BATCH * (hi - lo) * DIM
    tf_ach = flops / (kern_ms * 1e-3) / 1e12
    fmt = {L.SIM_BF16: "bf16", L.SIM_BF16_REFINE: "bf16", L.SIM_F16: "fp16", L.SIM_F16_REFINE: "fp16"}.get(mode)
    if main_run["filter"] is not None and fmt is not None:
        fmt = main_run["filter"]["filter"]          # automatic mode: what the store's policy settled on during warm-up
    roof = {"bound": "tensor", "achieved": tf_ach, "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": tf_ach / peaks["bf16"],
            "traffic": _ncu_traffic(f"cosine_topk_ts_kernel:N={hi - lo}:d={DIM}:Q={Q_BATCH}") if fmt else None,
            "kernel": (f"cosine_topk_ts_kernel (tcgen05 {fmt}, query tile stationary in TMEM) + threshold pre-pass"
                       + (" + fp32 refine (+ second tensor-core pass for uncertified rows)" if exact else "")) if fmt
            else "cosine_topk_f32_kernel (CUDA-core fp32)",
            "kernel_ms": kern_ms, "peak_source": peaks["src"] + " bf16 burst (cuBLAS 8192^3; fp16 runs at the same tensor rate)",
            "peak_sustained": peaks["bf16_sustained"], "frac_of_sustained": tf_ach / peaks["bf16_sustained"],
            "algorithmic": "2*Q*N_local*d flop per launch", "kernel_share_of_step": kern_ms / ms_step}

    line = {"metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None,
            "dtype": (f"{fmt} filter + f32 refine (exact)" if exact else fmt) if fmt else "f32",
            "data": "synthetic",
            "config": {"workload": f"top-{TOPK} cosine retrieve + value/label gather, {N_KEYS} keys d={DIM} sharded by key rows over "
                                   f"{world} GPU(s), {Q_BATCH}-query batches", "mode": mode, "data": args.data,
                       "filter": main_run["filter"], "filter_stable_during_timing": main_run["filter_stable"],
                       "l2": "inputs larger than L2 (key shard streamed every step)", "parallelism": f"key-row shard x{world}",
                       "exchange": {"p2p": "one kernel over NVLink peer memory (push candidates, merge, owners store rows to peers)",
                                    "nccl": "NCCL all-gather + merge + owner gather + all-gather", "single": "none"}[main_run["exchange"]]},
            "roofline": roof, "parity": main_run.get("parity"),
            "e2e": main_run["e2e"],
            "gpu_launches": main_run["launches"], "launches_per_step": main_run["launches"] / args.steps,
            "clocks": main_run["clocks"]}

    # ---- realistic library distributions: same shard sizes, keys regenerated in place ---------------
    if not args.no_extras:
        variants = {}
        for kind in ("clustered", "dup5"):
            if kind == args.data:
                continue
            fill_library_shard(store, lo, hi, DIM, N_CLASS, dev, kind)
            r = run_config(kind, max(3, min(5, args.steps)), 3, min(128, args.parity_rows), 16, False)
            variants[kind] = {"value": r["qps"], "unit": "queries/s", "ms_per_step": r["ms_per_step"], "filter": r["filter"],
                              "vs_gauss": r["qps"] / qps, "parity": r.get("parity"),
                              "pass2_row_fraction": (r["parity"]["pass2_rows"] / (Q_BATCH * world)) if r.get("parity") else None,
                              "data": {"clustered": f"{N_CENTROIDS} Gaussian centroids, sigma {CLUSTER_SIGMA}; queries = perturbed members",
                                       "dup5": "clustered + 5 % exact duplicate rows"}[kind]}
        line["variants"] = variants

    # ---- everything below is single-GPU work on rank 0 (replicas only); the library is released first ---
    del sr
    store.clear(release=True)
    del store
    torch.cuda.empty_cache()
    if rank == 0:
        if not args.no_spmm:
            try:
                line["spmm"] = bench_spmm(dev, args, peaks, stock=(world == 1 and not args.no_extras))
            except torch.cuda.OutOfMemoryError as e:
                line["spmm"] = {"skipped": f"OOM: {str(e)[:80]}"}
        if world == 1 and not args.no_extras:
            for name, fn in (("gather", bench_gather), ("cfg3", bench_cfg3), ("gpu_stock_baseline", bench_stock_retrieve),
                             ("small", bench_small), ("edge_widek", bench_edge_widek), ("edge_eval", bench_edge_eval)):
                try:
                    line[name] = fn(dev, args, peaks)
                except Exception as e:                              # an extra must never cost the headline line
                    line[name] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
                torch.cuda.empty_cache()
            if isinstance(line.get("gpu_stock_baseline"), dict) and "qps_extrapolated" in line["gpu_stock_baseline"]:
                line["gpu_stock_baseline"]["ours_over_stock"] = qps / line["gpu_stock_baseline"]["qps_extrapolated"]
        if not args.no_cpu_baseline and world == 1:
            t_full, desc, cores, _ = cpu_fit([args.cpu_sample // 4, args.cpu_sample], 3, 1)
            line["cpu_baseline"] = {"value": Q_BATCH / t_full, "unit": "queries/s", "cores": cores, "kind": "port", "sample": desc}
        line["bench_wall_s"] = time.perf_counter() - t_wall0
    if world > 1:
        dist.barrier()
    guard.restore()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        guard = StdoutToStderr()            # teardown chatter stays off stdout too
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------- single-GPU sections
def bench_spmm(dev, args, peaks, stock=True):
    from ragraph_b200 import _lib as L, ops
    rowptr, col, val, max_deg = make_products_graph(dev)
    x = torch.randn(SPMM_N, SPMM_F, device=dev)
    steps = max(args.steps, 5)
    l0 = L.launch_count()
    ms = timeit_events(lambda: ops.csr_spmm(rowptr, col, val, x), steps)
    launches = int(L.launch_count() - l0)
    # property check at full size: A.1 == row sums of val
    ones = ops.csr_spmm(rowptr, col, val, torch.ones(SPMM_N, 16, device=dev))[:, 0].double()
    rs = torch.zeros(SPMM_N, device=dev, dtype=torch.float64).index_add_(
        0, torch.repeat_interleave(torch.arange(SPMM_N, device=dev), rowptr[1:] - rowptr[:-1]), val.double())
    assert float((ones - rs).abs().max() / rs.abs().max()) < 1e-5, "SpMM row-sum property failed"
    # parity at full size on sampled rows: fp64 evaluation of those rows
    y = ops.csr_spmm(rowptr, col, val, x)
    worst = 0.0
    rp = rowptr.cpu()
    for r in range(0, SPMM_N, SPMM_N // 256):
        a, b = int(rp[r]), int(rp[r + 1])
        if b == a:
            continue
        ref = (val[a:b].double()[:, None] * x[col[a:b].long()].double()).sum(0)
        worst = max(worst, float((y[r].double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30)))
    assert worst < 1e-5, f"SpMM sampled-row parity failed: {worst}"
    alg = SPMM_NNZ * 8 + (SPMM_N + 1) * 8 + SPMM_NNZ * SPMM_F * 4 + SPMM_N * SPMM_F * 4
    gbs = alg / (ms * 1e-3) / 1e9
    out = {"value": gbs, "unit": "GB/s (algorithmic, edge-gather model)", "ms": ms, "edges_per_s": SPMM_NNZ / (ms * 1e-3),
           "config": {"workload": f"CSR SpMM ogbn-products-shaped synthetic: n={SPMM_N} nnz={SPMM_NNZ} F={SPMM_F} fp32, "
                                  f"Chung-Lu power law, max degree {max_deg}", "l2": "X (2.5 GB) larger than L2"},
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                        "traffic": _ncu_traffic(f"csr_spmm_kernel:n={SPMM_N}:nnz={SPMM_NNZ}:F={SPMM_F}"),
                        "kernel": "csr_spmm_kernel<32,2>", "algorithmic_bytes": alg,
                        "compulsory_bytes": SPMM_NNZ * 8 + (SPMM_N + 1) * 8 + 2 * SPMM_N * SPMM_F * 4,
                        "peak_source": peaks["src"] + " copy bandwidth"},
           "parity": {"rows_checked_fp64": 256, "max_rel_err": worst, "row_sum_property": "ok"},
           "gpu_launches": launches}
    if stock:
        # the reference's formulations on stock torch CUDA: torch.sparse.mm (cuSPARSE CSR) and the edge variant's
        # x[src] * w -> scatter_add_ (modules/RAGraph.py:232-240), edge-chunked because the [E, d] temp is 63 GB
        try:
            A = torch.sparse_csr_tensor(rowptr, col.long(), val, size=(SPMM_N, SPMM_N))
            ms_s = timeit_events(lambda: torch.sparse.mm(A, x), 3, 1)
            yd = torch.sparse.mm(A, x)
            out["stock_torch_sparse_mm"] = {"ms": ms_s, "edges_per_s": SPMM_NNZ / (ms_s * 1e-3), "ours_speedup": ms_s / ms,
                                            "max_rel_diff_vs_ours": float((yd - y).abs().max() / y.abs().max())}
            del A, yd
            dst = torch.repeat_interleave(torch.arange(SPMM_N, device=dev), rowptr[1:] - rowptr[:-1])
            src = col.long()
            ch = 8_000_000

            def ref_agg():
                o = torch.zeros(SPMM_N, SPMM_F, device=dev)
                for a in range(0, SPMM_NNZ, ch):
                    b = min(SPMM_NNZ, a + ch)
                    o.scatter_add_(0, dst[a:b, None].expand(-1, SPMM_F), x[src[a:b]] * val[a:b, None])
                return o
            ms_a = timeit_events(ref_agg, 2, 1)
            out["stock_torch_scatter_add"] = {"ms": ms_a, "edges_per_s": SPMM_NNZ / (ms_a * 1e-3), "ours_speedup": ms_a / ms,
                                              "note": "edge-chunked (8 M edges) only because the [E, d] temporary does not fit"}
        except Exception as e:
            out["stock_error"] = f"{type(e).__name__}: {str(e)[:120]}"
    return out


def bench_gather(dev, args, peaks):
    """retrieved-subgraph gather, ogbn-products shape (BASELINE config 5, second half): Q = 2.4 M, k = 10, d = 256"""
    from ragraph_b200 import _lib as L, ops
    N, d, Q, k = SPMM_N, 256, 2_400_000, 10
    table = torch.randn(N, d, device=dev)
    idx = torch.randint(0, N, (Q, k), device=dev)
    ms = timeit_events(lambda: ops.gather_rows(table, idx), 5, 2)
    alg = Q * k * (8 + 2 * d * 4)
    ms_t = timeit_events(lambda: table[idx], 3, 1)
    ms_r = timeit_events(lambda: ops.gather_reduce(table, idx, L.REDUCE_MEAN), 5, 2)
    alg_r = Q * k * (8 + d * 4) + Q * d * 4
    ms_rt = timeit_events(lambda: table[idx].mean(1), 3, 1)
    ok = bool(torch.equal(ops.gather_rows(table, idx[:200000]), table[idx[:200000]]))
    assert ok, "gather not bit exact"
    return {"workload": f"values[idx]: table {N} x {d} fp32, idx [{Q}, {k}] uniform",
            "gather_rows": {"ms": ms, "gbs": alg / ms / 1e6, "frac_of_hbm": alg / ms / 1e6 / peaks["hbm"], "algorithmic_bytes": alg,
                            "stock_torch_index_ms": ms_t, "ours_speedup": ms_t / ms, "bit_exact": ok},
            "gather_reduce_mean": {"ms": ms_r, "gbs": alg_r / ms_r / 1e6, "frac_of_hbm": alg_r / ms_r / 1e6 / peaks["hbm"],
                                   "algorithmic_bytes": alg_r, "stock_torch_index_mean_ms": ms_rt, "ours_speedup": ms_rt / ms_r}}


def bench_cfg3(dev, args, peaks):
    """BASELINE config 3: 10 M keys x d = 256, 4 096-query batches, top-10, one GPU; Gaussian and clustered keys"""
    from ragraph_b200 import _lib as L, ops
    N, d = 10_000_000, 256
    out = {"workload": f"top-{TOPK} cosine, {N} keys d={d}, Q={Q_BATCH}, exact mode"}
    store = make_library_shard(0, N, d, N_CLASS, dev, "gauss")
    for kind in ("gauss", "clustered"):
        if kind != "gauss":
            fill_library_shard(store, 0, N, d, N_CLASS, dev, kind)
        q = make_queries(Q_BATCH, d, dev, kind=kind).to(dev)
        for _ in range(4):                                       # lets the filter-format policy settle (untimed)
            store.topk_local(q, TOPK)
            torch.cuda.synchronize()
        store.collect_stats = True
        s, i = store.topk_local(q, TOPK)
        store.collect_stats = False
        st = store.last_stats.tolist()
        rows = torch.arange(0, Q_BATCH, 32, device=dev)
        s0, i0 = ops.cosine_topk(q[rows].contiguous(), store.resource_keys, TOPK, store.key_inv_norm)
        differ = (i[rows] != i0).any(dim=1)
        tie_ok = bool(((s[rows] - s0).abs().max(dim=1).values[differ] < 1e-6).all()) if bool(differ.any()) else True
        assert float((s[rows] - s0).abs().max()) < 2e-6 and tie_ok, "cfg3 parity vs the fp32 kernel failed"
        state = store.auto_state()
        ms = timeit_events(lambda: store.topk_local(q, TOPK), max(5, args.steps), 2)
        tf = 2.0 * Q_BATCH * N * d / ms / 1e9
        out[kind] = {"ms": ms, "qps": Q_BATCH / ms * 1e3, "tflops": tf, "frac_of_peak": tf / peaks["bf16"], "filter": state,
                     "pass2_rows": int(st[0]), "fallback_rows": int(st[1]),
                     "parity": {"rows": int(rows.numel()), "vs": "fp32 kernel", "rows_with_tie_swaps": int(differ.sum()),
                                "max_score_diff": float((s[rows] - s0).abs().max())}}
    store.clear(release=True)
    return out


def bench_stock_retrieve(dev, args, peaks):
    """The kernel to beat on the same box: the reference's retrieve on stock torch CUDA (fp32 sgemm + at::topk + index), timed on
    a bounded key sample and extrapolated linearly in N (its cost is one sgemm + one topk per chunk)."""
    n_s, chunk = 4_194_304, 262_144
    g = torch.Generator(device=dev).manual_seed(1234)
    keys = F.normalize(torch.randn(n_s, DIM, generator=g, device=dev), dim=-1)
    vals = torch.randn(n_s, DIM, generator=g, device=dev)
    q = make_queries(Q_BATCH, DIM, dev).to(dev)

    def ref():
        s, i = stock_topk(q, keys, TOPK, chunk)
        return vals[i]
    ms = timeit_events(ref, 3, 1)
    ms_full = ms * (N_KEYS / n_s)
    return {"what": "F.normalize + matmul + torch.topk + index (SimilarityFunctions.py:6-16, ToyGraphBase.py:67-71) on stock torch CUDA, "
                    f"fp32 (allow_tf32 off), key chunks of {chunk}", "sample_keys": n_s, "ms_at_sample": ms,
            "qps_extrapolated": Q_BATCH / (ms_full * 1e-3), "ms_per_step_extrapolated": ms_full,
            "note": f"timed on {n_s} keys, extrapolated linearly to {N_KEYS}"}


def bench_small(dev, args, peaks):
    """The reference's real shapes (BASELINE configs 1 and 2): cfg1 = Cora-shaped node batch (Q = 2 708 nodes against a
    10 832-row library, d = 256, k = 4), cfg2 = one graph-level query against a 480-row library (d = 256, k = 3) -- ours vs
    the reference's torch calls on the same GPU, wall time per call including host overhead (what a training loop sees)."""
    import ragraph_b200 as R

    def wall_us(fn, iters):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / iters * 1e6

    out = {}
    for name, Q, N, d, C, k, variant in (("cfg1", 2708, 10832, 256, 3, 4, "node"), ("cfg2", 1, 480, 256, 6, 3, "graph")):
        g = torch.Generator(device=dev).manual_seed(5)
        keys = F.normalize(torch.randn(N, d, generator=g, device=dev), dim=-1)
        vals = torch.randn(N, d, generator=g, device=dev)
        labs = F.one_hot(torch.randint(0, C, (N,), generator=g, device=dev), C).float()
        q = torch.randn(Q, d, generator=g, device=dev)
        base = R.ToyGraphBase(None, C, d, 3, device=dev, variant=variant, capacity=N)
        base.retrieve_num = k
        base.add_entries(keys, vals, labs)
        qq = q[0] if variant == "graph" else q

        def ours():
            return base.retrieve(qq, None, False)

        def stock():
            s = torch.matmul(F.normalize(q, p=2, dim=-1), F.normalize(keys, p=2, dim=-1).t())
            _, i = torch.topk(s, k, largest=True, sorted=True)
            return vals[i], labs[i]
        e1, l1 = ours(); e0, l0 = stock()
        rows_same = (e1 == e0).flatten(1).all(dim=1)
        same = float(rows_same.float().mean())                  # rows that differ are near-ties of the fp32 scores (checked in tests)
        iters = 200 if Q > 1 else 2000
        out[name] = {"workload": f"retrieve Q={Q} N={N} d={d} k={k}", "ours_us": wall_us(ours, iters),
                     "stock_torch_us": wall_us(stock, iters), "rows_identical_to_stock": same}
        out[name]["ours_speedup"] = out[name]["stock_torch_us"] / out[name]["ours_us"]
    out["cfg1_forward"] = _bench_forward(dev, wall_us, "node", 2708, 1433, 256, 7, 10832, 4, 3, 0.5)
    out["cfg2_forward"] = _bench_forward(dev, wall_us, "graph", 40, 3, 256, 2, 480, 3, 1, 0.3)
    return out


def _bench_forward(dev, wall_us, variant, n, f, d, C, N, k, hops, w):
    """BASELINE configs 1 / 2 as the reference runs them: a whole RAGraph.forward (RAGraph_node/RAGraph.py:39-63 over a
    Cora-shaped graph; RAGraph_graph/RAGraph.py:58-75 over one PROTEINS-shaped graph, the query = mean node embedding) --
    retrieve + gathers + k-hop propagation + decoder + fusion, with a one-layer GCN as the pre-trained encoder.  Ours
    eager, ours replayed as one CUDA graph (graphs.GraphedForward), the reference's formulas on stock torch."""
    import ragraph_b200 as R
    g = torch.Generator(device=dev).manual_seed(3)
    a = (torch.rand(n, n, generator=g, device=dev) < 3.9 / n).float()
    a = torch.triu(a, 1); a = a + a.t() + torch.eye(n, device=dev)
    dinv = a.sum(1).pow(-0.5); adj = (dinv[:, None] * a * dinv[None, :]).contiguous()
    x = torch.randn(n, f, generator=g, device=dev)
    keys = F.normalize(torch.randn(N, d, generator=g, device=dev), dim=-1)
    vals = torch.randn(N, d, generator=g, device=dev)
    labs = F.one_hot(torch.randint(0, C, (N,), generator=g, device=dev), C).float()

    class Enc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gcn = R.GCN(f, d, "prelu")

        def inference(self, features, adj_):
            return self.gcn([features, adj_])
    base = R.ToyGraphBase(None, C, d, 3, device=dev, capacity=N, variant=variant)
    base.retrieve_num = k
    base.add_entries(keys, vals, labs)
    model = R.RAGraph(Enc(), base, f, C, d, variant=variant).to(dev).eval()
    enc, dec = model.pretrain_model.gcn, model.decoder

    def ours():
        with torch.no_grad():
            return model(x, adj)

    def stock():
        with torch.no_grad():
            emb = enc.act(torch.mm(adj, enc.fc(x)) + enc.bias)                      # layers/gcn.py:32-40
            qv = emb if variant == "node" else emb.mean(0, keepdim=True)
            s = torch.matmul(F.normalize(qv, p=2, dim=-1), F.normalize(keys, p=2, dim=-1).t())
            _, i = torch.topk(s, k, largest=True, sorted=True)
            re, rl = vals[i].sum(1), labs[i].mean(1)
            an = adj / adj.sum(1, keepdim=True)
            h = emb
            for _ in range(hops):
                h = F.relu(torch.matmul(an, h))
            if variant == "graph":
                h = h.mean(0, keepdim=True)
            h = h * (1 - w) + re * w
            return torch.softmax(dec(h), 1) * (1 - w) + rl * w
    graphed = R.GraphedForward(model, x, adj)
    o_e, o_s, o_g = ours(), stock(), graphed(x).clone()
    res = {"workload": f"RAGraph.forward {variant} variant: n={n} f={f} d={d} library N={N} k={k}, {hops} hop(s) (dense [n,n] adjacency in)",
           "ours_eager_us": wall_us(ours, 200), "ours_graphed_us": wall_us(lambda: graphed(x), 200),
           "stock_torch_us": wall_us(stock, 200), "graphed_equals_eager": bool(torch.equal(o_e, o_g)),
           "max_abs_diff_vs_stock": float((o_e - o_s).abs().max())}
    res["graphed_speedup_vs_stock"] = res["stock_torch_us"] / res["ours_graphed_us"]
    res["eager_speedup_vs_stock"] = res["stock_torch_us"] / res["ours_eager_us"]
    return res


def bench_edge_widek(dev, args, peaks):
    """The edge variant's vanilla phase (RAGraph_edge/modules/RAGraph.py:36-38, 298-311): batches of 32 768 node embeddings
    (d = 64) against the whole resource library, retrieve_num = 50 -- k beyond the 32-entry candidate lists, served on the
    tensor cores by MORE KEY SPLITS.  Ours (store.topk: automatic exact mode) vs the fp32 CUDA-core kernel and
    the reference's calls on stock torch (normalize + matmul + topk, query-chunked: [32768, 240000] fp32 is 31 GB)."""
    import ragraph_b200 as R
    from ragraph_b200 import _lib as L, ops
    Q, N, d, k = 32768, 240_000, 64, 50
    g = torch.Generator(device=dev).manual_seed(11)
    keys = torch.randn(N, d, generator=g, device=dev)
    q = torch.randn(Q, d, generator=g, device=dev)
    base = R.ToyGraphBase(None, 2, d, 3, device=dev, capacity=N)
    base.add_entries(keys, keys, torch.zeros(N, 2, device=dev))
    s1, i1 = base.topk(q, k)
    inv = ops.row_inv_norm(base.resource_keys)
    s0, i0 = ops.cosine_topk(q, base.resource_keys, k, inv)
    chunk = 4096

    def stock():
        kn = F.normalize(keys, p=2, dim=-1)
        outs = [torch.topk(torch.matmul(F.normalize(q[a:a + chunk], p=2, dim=-1), kn.t()), k, largest=True, sorted=True)
                for a in range(0, Q, chunk)]
        return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])
    ss, si = stock()
    ms = timeit_events(lambda: base.topk(q, k), 5, 2)
    ms_f32 = timeit_events(lambda: ops.cosine_topk(q, base.resource_keys, k, inv), 2, 1)
    ms_stock = timeit_events(stock, 2, 1)
    # rows whose index lists differ from the fp32 kernel's / stock torch's must be near-ties: compare the score rows
    return {"workload": f"top-{k} cosine, Q={Q} N={N} d={d} (edge variant, vanilla phase)", "ours_ms": ms,
            "fp32_kernel_ms": ms_f32, "stock_torch_ms": ms_stock, "ours_over_stock": ms_stock / ms, "ours_over_fp32_kernel": ms_f32 / ms,
            "tflops": 2.0 * Q * N * d / ms / 1e9, "max_score_diff_vs_fp32_kernel": float((s1 - s0).abs().max()),
            "max_score_diff_vs_stock": float((s1 - ss).abs().max()),
            "rows_idx_differ_vs_stock": int((i1 != si).any(dim=1).sum()), "rows_idx_differ_vs_fp32_kernel": int((i1 != i0).any(dim=1).sum())}


def bench_edge_eval(dev, args, peaks):
    """The edge variant's evaluation ranking (RAGraph_edge/utils/metrics.py:96-118): a batch of user embeddings against
    every item (d = 64), history items excluded, top-20.  Ours on the tensor cores (edge.rating_topk: fp16 filter over the
    scaled item table + fp32 refine that drops the history + certificate), ours on the fp32 CUDA-core kernel, and the
    reference's formulation on stock torch kept on the GPU (matmul -> masked_fill of the history -> topk; the reference
    moves the [B, n_items] matrix to the CPU first, :110-117)."""
    from ragraph_b200 import edge as E
    B, I, d, k, h = 4096, 100_000, 64, 20, 30
    g = torch.Generator(device=dev).manual_seed(13)
    users = torch.randn(B, d, generator=g, device=dev) * 0.3
    items = torch.randn(I, d, generator=g, device=dev) * (0.5 + torch.rand(I, 1, generator=g, device=dev))
    hist_items = torch.randint(0, I, (B, h), generator=g, device=dev).sort(dim=1).values.reshape(-1)
    rowptr = torch.arange(0, (B + 1) * h, h, device=dev, dtype=torch.int64)
    rows = torch.arange(B, device=dev).repeat_interleave(h)

    def stock():
        pred = torch.matmul(users, items.t())
        pred[rows, hist_items] = float("-inf")
        return torch.topk(pred, k)[1]
    i_tc = E.rating_topk(users, items, k, rowptr, hist_items, tensor_cores=True)
    i_f32 = E.rating_topk(users, items, k, rowptr, hist_items, tensor_cores=False)
    i_st = stock()
    ms_tc = timeit_events(lambda: E.rating_topk(users, items, k, rowptr, hist_items, tensor_cores=True), 10, 3)
    ms_f32 = timeit_events(lambda: E.rating_topk(users, items, k, rowptr, hist_items, tensor_cores=False), 5, 2)
    ms_st = timeit_events(stock, 5, 2)
    return {"workload": f"masked dot-product top-{k}: {B} users x {I} items, d={d}, {h} history items per user",
            "ours_tensor_core_ms": ms_tc, "ours_fp32_kernel_ms": ms_f32, "stock_torch_gpu_ms": ms_st,
            "ours_over_stock": ms_st / ms_tc, "rows_identical_tc_vs_fp32_kernel": float((i_tc == i_f32).all(dim=1).float().mean()),
            "rows_identical_tc_vs_stock": float((i_tc == i_st).all(dim=1).float().mean())}


if __name__ == "__main__":
    main()
