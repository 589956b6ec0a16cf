#!/usr/bin/env python
"""Device-timed numbers for the kernels added after the core path (one JSON line each, CUDA events, L2-exceeding inputs):
tf32 filter (TFLOP/s, recall@10 against the exact mode), prototype scores / prompt (GB/s), per-destination softmax (GB/s).

    python tools/bench_widen.py > gpurun_out/bench_widen.jsonl
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ragraph_b200 import _lib as L, ops  # noqa: E402

DEV = "cuda"


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2]


def emit(**kw):
    print(json.dumps(kw), flush=True)


def section(name, fn):
    t0 = time.time()
    try:
        fn()
    except Exception as e:  # keep going: every section is independent evidence
        emit(section=name, error=repr(e)[:300])
    emit(section=name, wall_s=round(time.time() - t0, 2))


def tf32():
    Q, N, k = 4096, 2_000_000, 10
    g = torch.Generator(device=DEV).manual_seed(1)
    for d in (128, 64):
        keys = torch.nn.functional.normalize(torch.randn(N, d, generator=g, device=DEV), dim=-1)
        q = torch.randn(Q, d, generator=g, device=DEV)
        inv = ops.row_inv_norm(keys)
        sh32, sh16 = ops.rows_to_tf32(keys), ops.rows_to_bf16(keys)
        L.tc_set_option("variant", 1)
        res = {}
        for name, mode, sh in (("tf32", L.SIM_TF32, sh32), ("bf16", L.SIM_BF16, sh16), ("exact", L.SIM_BF16_REFINE, sh16)):
            ms = timeit(lambda: ops.cosine_topk(q, keys, k, inv, sh, mode), 5, 2)
            res[name] = (ms, ops.cosine_topk(q, keys, k, inv, sh, mode))
        L.tc_set_option("variant", -1)
        ex = res["exact"][1][1]
        rec = {n: float((res[n][1][1].unsqueeze(2) == ex.unsqueeze(1)).any(2).float().mean()) for n in ("tf32", "bf16")}
        err = {n: float((res[n][1][0] - res["exact"][1][0]).abs().max()) for n in ("tf32", "bf16")}
        emit(section="tf32_filter", Q=Q, N=N, d=d, k=k, kernel="cosine_topk_tc_kernel (SS)",
             ms={n: round(res[n][0], 4) for n in res}, tflops={n: round(2.0 * Q * N * d / (res[n][0] * 1e-3) / 1e12, 1) for n in res},
             recall_at_10_vs_exact=rec, max_abs_score_diff_vs_exact_topk=err)
        del keys, sh32, sh16


def prompt():
    n, d, C = 4_000_000, 256, 7
    g = torch.Generator(device=DEV).manual_seed(2)
    x = torch.randn(n, d, generator=g, device=DEV)
    w = torch.randn(d, generator=g, device=DEV)
    proto = torch.randn(C, d, generator=g, device=DEV)
    ms = timeit(lambda: ops.prototype_scores(x, proto, L.SCORES_SOFTMAX, w, L.ACT_ELU))
    emit(section="prototype_scores", n=n, d=d, C=C, ms=round(ms, 4), algorithmic_bytes=n * d * 4 + n * C * 4,
         gbs=round((n * d * 4 + n * C * 4) / (ms * 1e-3) / 1e9, 1))
    ms = timeit(lambda: ops.prompt_act(x, w, L.ACT_ELU))
    emit(section="prompt_act", n=n, d=d, ms=round(ms, 4), algorithmic_bytes=2 * n * d * 4, gbs=round(2 * n * d * 4 / (ms * 1e-3) / 1e9, 1))


def softmax():
    E, n = 61_859_140, 2_449_029
    g = torch.Generator(device=DEV).manual_seed(3)
    idx = torch.randint(0, n, (E,), generator=g, device=DEV)
    t = torch.rand(E, generator=g, device=DEV)
    base = torch.rand(E, generator=g, device=DEV)
    ms = timeit(lambda: ops.scatter_softmax(t, idx, n, 0.0, 1.0, base, 0.5, 0.5))
    alg = 3 * E * 12 + E * 4 + E * 4         # three passes over (src, index) + base read + out write
    emit(section="scatter_softmax", E=E, n_groups=n, ms=round(ms, 4), algorithmic_bytes=alg, gbs=round(alg / (ms * 1e-3) / 1e9, 1),
         launches_per_call=4)


if __name__ == "__main__":
    section("prompt", prompt)
    section("softmax", softmax)
    section("tf32", tf32)
