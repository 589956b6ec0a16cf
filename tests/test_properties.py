"""Property-based CPU tests (hypothesis) of the host logic and of the oracle's own helpers: shard partition /
ownership, shard-merge == single scan (with exact ties across shards), CSR-direct preprocessing == the reference's dense
block-diagonal adjacency, deterministic COO -> CSR, the parity criterion itself.  No compute call touches the CUDA library."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import ragraph_oracle as O
from ragraph_b200 import process_graph_batch
from ragraph_b200.csr import CSRGraph
from ragraph_b200.sharded import owner_of, shard_bounds

FAST = settings(max_examples=30, deadline=None, derandomize=True, database=None)   # same examples on every run


@FAST
@given(n=st.integers(0, 5000), world=st.integers(1, 16))
def test_shard_bounds_and_owner_agree(n, world):
    spans = [shard_bounds(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(b == c for (_, b), (c, _) in zip(spans, spans[1:]))
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1 and sorted(sizes, reverse=True) == sizes     # the extra rows go to the first ranks
    if n:
        own = owner_of(torch.arange(n), n, world)
        for r, (a, b) in enumerate(spans):
            assert bool((own[a:b] == r).all())


@FAST
@given(seed=st.integers(0, 2 ** 31 - 1), world=st.integers(1, 6), k=st.integers(1, 12), dup=st.booleans())
def test_shard_merge_equals_single_scan(seed, world, k, dup):
    """Per-shard top-k with global indices, merged in the product's order (score desc, index asc), equals the top-k of
    the whole library -- including exact duplicate keys that land in different shards."""
    g = torch.Generator().manual_seed(seed)
    Q, N, d = 7, 60, 8
    q = torch.randn(Q, d, generator=g)
    keys = torch.randn(N, d, generator=g)
    if dup:
        keys[N - 1] = keys[0]
        keys[N // 2] = keys[1]
    S = O.cosine_similarity(q, keys)
    parts_s, parts_i = [], []
    for r in range(world):
        lo, hi = shard_bounds(N, world, r)
        kk = min(k, hi - lo)
        # the shard's top-k in the product's deterministic order (score desc, index asc); torch.topk leaves the order of
        # ties -- and WHICH of two tied keys survives at the cut -- unspecified
        loc = np.lexsort((np.arange(hi - lo)[None, :].repeat(Q, 0), -S[:, lo:hi].numpy().astype(np.float64)), axis=1)[:, :kk]
        i = torch.from_numpy(loc) + lo                         # global row ids
        s = torch.gather(S, 1, i)
        if kk < k:                                             # tiny shard: padded like ShardedRetriever._local
            s = torch.cat([s, s.new_full((Q, k - kk), -torch.finfo(torch.float32).max)], 1)
            i = torch.cat([i, i.new_full((Q, k - kk), -1)], 1)
        parts_s.append(s); parts_i.append(i)
    ms, mi = O.merge_topk(torch.stack(parts_s), torch.stack(parts_i), k)
    # reference order: score desc, index asc over the whole row
    order = np.lexsort((np.arange(N)[None, :].repeat(Q, 0), -S.numpy().astype(np.float64)), axis=1)[:, :k]
    assert np.array_equal(mi.numpy(), order)
    assert np.array_equal(ms.numpy(), np.take_along_axis(S.numpy(), order, axis=1))
    ok, bad = O.topk_sets_match(mi.numpy(), O.cosine_similarity_f64(q, keys), k)
    assert ok, bad


@FAST
@given(seed=st.integers(0, 2 ** 31 - 1), k=st.integers(1, 6))
def test_parity_criterion_accepts_tie_swaps_and_rejects_wrong_rows(seed, k):
    rng = np.random.default_rng(seed)
    Q, N = 4, 30
    S = rng.standard_normal((Q, N))
    S[:, 7] = S[:, 3]                                          # an exact tie in every row
    top = np.argsort(-S, axis=1, kind="stable")[:, :k]
    assert O.topk_sets_match(top, S, k)[0]
    swapped = top.copy()
    for r in range(Q):                                         # exchanging the tied pair keeps the answer admissible
        swapped[r] = [7 if x == 3 else (3 if x == 7 else x) for x in swapped[r]]
    assert O.topk_sets_match(swapped, S, k)[0]
    worst = np.argsort(S, axis=1)[:, :1]                       # the lowest-scoring key can never be in the top k < N
    wrong = top.copy(); wrong[:, -1] = worst[:, 0]
    if k < N and not np.any(np.abs(S[np.arange(Q), worst[:, 0]] - np.sort(S, axis=1)[:, -k]) <= 1e-6):
        assert not O.topk_sets_match(wrong, S, k)[0]
    dupl = top.copy()
    if k >= 2:
        dupl[:, 1] = dupl[:, 0]
        assert not O.topk_sets_match(dupl, S, k)[0]


@FAST
@given(seed=st.integers(0, 2 ** 31 - 1), n_graphs=st.integers(1, 4))
def test_csr_preprocessing_equals_reference_dense_adjacency(seed, n_graphs):
    g = torch.Generator().manual_seed(seed)
    xs, eis = [], []
    for _ in range(n_graphs):
        n = int(torch.randint(1, 12, (1,), generator=g))
        E = int(torch.randint(0, 3 * n, (1,), generator=g))
        xs.append(torch.rand(n, 5, generator=g))
        eis.append(torch.randint(0, n, (2, E), generator=g))
    feats, csr, labs = process_graph_batch(xs, eis, 3)
    rf, radj, rl = O.process_tu_arrays([x.numpy() for x in xs], [e.numpy() for e in eis], 3)
    assert torch.equal(feats, rf) and torch.equal(labs, rl)
    dense = torch.zeros(csr.n_rows, csr.n_cols)
    dense[csr.row_ids(), csr.col.long()] = csr.val
    assert float((dense - radj).abs().max()) <= 1e-7 and torch.equal(dense != 0, radj != 0)
    assert int(csr.rowptr[-1]) == csr.nnz and bool((csr.rowptr[1:] >= csr.rowptr[:-1]).all())


@FAST
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 40), E=st.integers(0, 200), weighted=st.booleans())
def test_deterministic_coo_to_csr_is_the_same_matrix(seed, n, E, weighted):
    g = torch.Generator().manual_seed(seed)
    edges = torch.stack([torch.randint(0, n, (E,), generator=g), torch.randint(0, n, (E,), generator=g)], 1)
    w = torch.rand(E, generator=g) if weighted else None
    csr = CSRGraph.from_coo(edges, w, n, n, deterministic=True)
    X = torch.randn(n, 3, generator=g).double()
    ref = torch.from_numpy(O.edge_agg_f64(X.numpy(), edges.numpy(), (w if w is not None else torch.ones(E)).numpy(), n))
    val = csr.val.double() if csr.val is not None else torch.ones(E, dtype=torch.float64)
    got = torch.zeros(n, 3, dtype=torch.float64).index_add_(0, csr.row_ids(), X[csr.col.long()] * val[:, None])
    assert torch.allclose(got, ref, atol=1e-9)
    # inside a row, entries keep the original edge order (fixed fp32 summation order)
    t = csr.transpose().transpose()
    assert t is csr


def _round_tf32_rna(x: torch.Tensor) -> torch.Tensor:
    """cvt.rna.tf32.f32 on the CPU: keep 10 mantissa bits, round to nearest, ties away from zero."""
    bits = x.contiguous().view(torch.int32)
    mag = (bits & 0x7fffffff) + 0x1000                      # half of the dropped 13 bits: ties go away from zero
    return ((bits & -0x80000000) | (mag & ~0x1fff)).view(torch.float32)


@FAST
@given(seed=st.integers(0, 2 ** 31 - 1), d=st.sampled_from([8, 30, 64, 128, 256]),
       kind=st.sampled_from(["gauss", "sparse", "constant", "heavy"]))
def test_low_precision_score_error_bounds(seed, d, kind):
    """The premise of the exactness certificate (csrc/topk_tc.cu, TC_EPS) and of the tf32 error band the tests use: for
    L2-normalised vectors whose components are rounded once to bf16 (8 significant bits) or tf32 (11), the dot product of
    the rounded vectors differs from the exact cosine by at most 2^-8 resp. 2^-10 (plus fp32 accumulation slack)."""
    g = torch.Generator().manual_seed(seed)
    n = 64
    if kind == "gauss":
        q, k = torch.randn(n, d, generator=g), torch.randn(n, d, generator=g)
    elif kind == "sparse":                                  # almost one-hot: a single large component
        q = torch.randn(n, d, generator=g) * 1e-3; k = torch.randn(n, d, generator=g) * 1e-3
        q[torch.arange(n), torch.randint(0, d, (n,), generator=g)] = 1.0
        k[torch.arange(n), torch.randint(0, d, (n,), generator=g)] = -1.0
    elif kind == "constant":                                # all components equal: every rounding error has the same sign
        q = torch.full((n, d), 1.0) * torch.rand(n, 1, generator=g); k = q.clone()
    else:                                                   # heavy tailed
        q = torch.randn(n, d, generator=g) ** 3; k = torch.randn(n, d, generator=g) ** 3
    qn, kn = torch.nn.functional.normalize(q, dim=-1), torch.nn.functional.normalize(k, dim=-1)
    exact = (qn.double() * kn.double()).sum(-1)
    bf = (qn.bfloat16().double() * kn.bfloat16().double()).sum(-1)
    assert float((bf - exact).abs().max()) <= 2.0 ** -8 + 2.0 ** -10
    tf = (_round_tf32_rna(qn).double() * _round_tf32_rna(kn).double()).sum(-1)
    assert float((tf - exact).abs().max()) <= 2.0 ** -10 + 1e-5
    r = _round_tf32_rna(qn)
    assert int((r.view(torch.int32) & 0x1fff).abs().max()) == 0 and float((r - qn).abs().max()) <= 2.0 ** -11


def _filter_refine_model(q, keys, k, kp, n_splits, eps, thr0=None):
    """NumPy model of the exact mode's ALGORITHM, independent of the CUDA code: per key split keep the
    kp best bf16 scores (only scores above the optional pre-pass bound thr0 are ever listed), then per row: tau = k-th best
    approximate score, re-score exactly every candidate with approximate score >= tau - 2 eps, certify iff the k-th exact
    score beats max(split thresholds, thr0) + eps.  Returns (idx [Q,k], certified [Q])."""
    qn = torch.nn.functional.normalize(q, dim=-1); kn = torch.nn.functional.normalize(keys, dim=-1)
    approx = (qn.bfloat16().double() @ kn.bfloat16().double().T).numpy()
    exact = (qn.double() @ kn.double().T).numpy()
    Q, N = approx.shape
    bounds = np.linspace(0, N, n_splits + 1).astype(int)
    out_idx = np.full((Q, k), -1, dtype=np.int64); cert = np.zeros(Q, dtype=bool)
    for r in range(Q):
        cand, tmax = [], -np.inf if thr0 is None else thr0[r]
        for s in range(n_splits):
            lo, hi = bounds[s], bounds[s + 1]
            sc = approx[r, lo:hi]
            idx = np.arange(lo, hi)
            if thr0 is not None:
                keep = sc > thr0[r]
                sc, idx = sc[keep], idx[keep]
            order = np.argsort(-sc, kind="stable")[:kp]
            cand.extend(idx[order].tolist())
            if len(order) == kp:                               # a full list: everything not listed scores <= its minimum
                tmax = max(tmax, sc[order[-1]])
        cand = np.array(sorted(set(cand)), dtype=np.int64)
        if cand.size < k:
            continue
        a = approx[r, cand]
        tau = np.sort(a)[-k]
        sel = cand[a >= tau - 2 * eps]
        e = exact[r, sel]
        order = np.lexsort((sel, -e))[:k]
        out_idx[r] = sel[order]
        cert[r] = e[order[-1]] > tmax + eps
    return out_idx, cert, exact


@FAST
@given(seed=st.integers(0, 2 ** 31 - 1), k=st.integers(1, 10), clustered=st.booleans(), prepass=st.booleans())
def test_filter_refine_certificate_model(seed, k, clustered, prepass):
    """Whenever the certificate holds, the refined answer IS the exact top-k (ties within 1e-6 aside) -- including dense
    clusters / exact duplicates at the boundary and rows started at a pre-pass bound; uncertified rows are exactly the ones
    the product recomputes in fp32."""
    g = torch.Generator().manual_seed(seed)
    Q, N, d, kp, n_splits = 6, 400, 32, 16, 3
    keys = torch.randn(N, d, generator=g)
    if clustered:
        cent = torch.randn(4, d, generator=g)
        keys = cent[torch.randint(0, 4, (N,), generator=g)] + 0.02 * keys
        keys[1] = keys[0]
    q = torch.randn(Q, d, generator=g)
    eps = 2.0 ** -8 + 2.0 ** -10
    thr0 = None
    if prepass:                                               # a valid lower bound of the kp-th best approximate score
        qn = torch.nn.functional.normalize(q, dim=-1); kn = torch.nn.functional.normalize(keys, dim=-1)
        ap = (qn.bfloat16().double() @ kn.bfloat16().double().T).numpy()
        thr0 = np.sort(ap[:, ::7], axis=1)[:, -kp]            # kp-th largest of a sample of distinct keys
    idx, cert, exact = _filter_refine_model(q, keys, k, kp, n_splits, eps, thr0)
    if cert.any():
        ok, bad = O.topk_sets_match(idx[cert], exact[cert], k)
        assert ok, bad
    if not clustered:
        assert cert.all()                                     # well separated random keys always certify


def _streaming_shared_threshold_model(approx_row, exact_row, k, kp, n_splits, eps, rng, margin_mult=2.0):
    """NumPy model of ONE query row through the query-stationary filter with cross-split threshold sharing (DESIGN.md 3.1,
    TcArgs::pool), independent of the CUDA code.  The key splits advance in an arbitrary interleaving; a split accepts a key
    iff its approximate score beats the split's threshold, keeps its kp best (threshold = the kp-th best once it has kp),
    and mirrors accepted scores into a pool of published scores (any subset may be visible, slots may be overwritten by later
    candidates: both are modelled by publishing each accepted score with probability 1/2 and dropping a random published
    one now and then).  At random moments the sweep takes the k-th largest published score minus margin_mult * eps as the
    shared bound, which every split folds into its threshold at a random later moment.  Refine: t_split = max(the threshold
    a split ended with, its list minimum if full); certify iff the k-th exact score > max t_split + eps."""
    N = approx_row.shape[0]
    bounds = np.linspace(0, N, n_splits + 1).astype(int)
    pos = bounds[:-1].copy()
    lists = [[] for _ in range(n_splits)]                  # (approx, idx)
    thr = np.full(n_splits, -np.inf)
    seen_bound = np.full(n_splits, -np.inf)
    published, shared = [], -np.inf
    live = [s for s in range(n_splits) if pos[s] < bounds[s + 1]]
    while live:
        s = live[rng.integers(len(live))]
        j = pos[s]; pos[s] += 1
        if pos[s] >= bounds[s + 1]:
            live.remove(s)
        if rng.random() < 0.3:                              # the split reads the shared bound (possibly a stale one)
            seen_bound[s] = shared
            thr[s] = max(thr[s], seen_bound[s])
        a = approx_row[j]
        if a > thr[s]:
            lists[s].append((a, j))
            if rng.random() < 0.5:
                published.append(a)
            if len(lists[s]) > kp:
                lists[s].sort(key=lambda t: (-t[0], t[1]))
                lists[s] = lists[s][:kp]
            if len(lists[s]) == kp:
                thr[s] = max(thr[s], min(t[0] for t in lists[s]))
        if published and rng.random() < 0.05:
            published.pop(rng.integers(len(published)))     # a slot overwritten before the sweep saw it
        if len(published) >= k and rng.random() < 0.2:      # a sweep
            shared = max(shared, np.sort(published)[-k] - margin_mult * eps - 1e-6)
    cand = sorted({j for l in lists for _, j in l})
    if len(cand) < k:
        return None, False
    tmax = max(max(thr[s], min((t[0] for t in lists[s]), default=-np.inf) if len(lists[s]) == kp else -np.inf) for s in range(n_splits))
    cand = np.array(cand, dtype=np.int64)
    a = approx_row[cand]
    tau = np.sort(a)[-k]
    sel = cand[a >= tau - 2 * eps]
    e = exact_row[sel]
    order = np.lexsort((sel, -e))[:k]
    return sel[order], bool(e[order[-1]] > tmax + eps)


@FAST
@given(seed=st.integers(0, 2 ** 31 - 1), k=st.integers(1, 10), clustered=st.booleans())
def test_cross_split_shared_threshold_model(seed, k, clustered):
    """Sharing thresholds across key splits never costs exactness: whatever subset of the candidates the sweep sees and
    however stale the bound a split folds in, a certified row is the exact top-k -- and with margin 2 eps rows whose splits
    never overflow a list do certify (the bound sits far enough under the k-th best by construction)."""
    g = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    Q, N, d, kp, n_splits = 4, 360, 32, 16, 4
    keys = torch.randn(N, d, generator=g)
    if clustered:
        cent = torch.randn(4, d, generator=g)
        keys = cent[torch.randint(0, 4, (N,), generator=g)] + 0.02 * keys
        keys[1] = keys[0]
    q = torch.randn(Q, d, generator=g)
    qn = torch.nn.functional.normalize(q, dim=-1); kn = torch.nn.functional.normalize(keys, dim=-1)
    approx = (qn.bfloat16().double() @ kn.bfloat16().double().T).numpy()
    exact = (qn.double() @ kn.double().T).numpy()
    eps = 2.0 ** -8 + 2.0 ** -10
    n_cert = 0
    for r in range(Q):
        idx, cert = _streaming_shared_threshold_model(approx[r], exact[r], k, kp, n_splits, eps, rng)
        if cert:
            n_cert += 1
            ok, bad = O.topk_sets_match(idx[None, :], exact[r][None, :], k)
            assert ok, bad
    if not clustered:
        assert n_cert > 0                                   # spread-out keys: the certificate does hold with a shared bound


@FAST
@given(seed=st.integers(0, 2 ** 31 - 1), k=st.integers(1, 24), n_groups=st.integers(1, 64), clustered=st.booleans())
def test_two_pass_collect_threshold_model(seed, k, n_groups, clustered):
    """Two-pass mode, independent of the CUDA code: the k-th largest of a row's group maxima (16-bit scores
    of distinct keys) minus 2 eps is a COLLECT threshold -- every member of the exact top k scores above it in 16 bits, so
    re-scoring what the collect pass returns gives the exact top k.  With fewer than k groups the bound is -inf (everything
    is collected).  Also pins the retry bound for rows whose collect area overflowed: the exact k-th score of ANY subset of
    at least k keys, minus one eps, still collects the whole exact top k."""
    g = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    Q, N, d = 4, 500, 32
    keys = torch.randn(N, d, generator=g)
    if clustered:
        cent = torch.randn(3, d, generator=g)
        keys = cent[torch.randint(0, 3, (N,), generator=g)] + 0.02 * keys
        keys[1] = keys[0]
    q = torch.randn(Q, d, generator=g)
    qn = torch.nn.functional.normalize(q, dim=-1); kn = torch.nn.functional.normalize(keys, dim=-1)
    approx = (qn.half().double() @ kn.half().double().T).numpy()
    exact = (qn.double() @ kn.double().T).numpy()
    eps = 2 * 2.0 ** -11 + 2.0 ** -15                        # worst-case fp16 rounding of two unit vectors + slack
    assert np.abs(approx - exact).max() <= eps
    bounds = np.linspace(0, N, n_groups + 1).astype(int)
    for r in range(Q):
        gmax = np.array([approx[r, a:b].max() if b > a else -np.inf for a, b in zip(bounds[:-1], bounds[1:])])
        kth = np.sort(gmax)[-k] if n_groups >= k else -np.inf
        thr = kth - 2 * eps - 1e-6 if np.isfinite(kth) else -np.inf
        collected = np.nonzero(approx[r] > thr)[0]
        assert collected.size >= min(k, N)
        e = exact[r, collected]
        got = collected[np.lexsort((collected, -e))[:k]]
        ok, bad = O.topk_sets_match(got[None, :], exact[r][None, :], k)
        assert ok, bad
        # retry bound: exact k-th score of an arbitrary subset that holds at least k collected keys
        if collected.size > k:
            sub = rng.choice(collected, size=int(rng.integers(k, collected.size + 1)), replace=False)
            thr2 = np.sort(exact[r, sub])[-k] - eps - 1e-6
            coll2 = np.nonzero(approx[r] > thr2)[0]
            e2 = exact[r, coll2]
            got2 = coll2[np.lexsort((coll2, -e2))[:k]]
            ok, bad = O.topk_sets_match(got2[None, :], exact[r][None, :], k)
            assert ok, bad


@FAST
@given(seed=st.integers(0, 2 ** 31 - 1), k=st.integers(1, 8), heavy=st.booleans())
def test_masked_refine_certificate_model(seed, k, heavy):
    """Exclusion lists on the tensor-core path: the filter ignores them (per-split lists hold the kp best
    keys, excluded or not), refine drops excluded candidates, and the UNCHANGED certificate -- k-th exact admissible score >
    every full list's minimum + eps -- still proves the answer: whatever the lists did not hold scores lower, admissible or
    not.  ``heavy``: the exclusions ARE the row's best keys, so few rows certify; the ones that do are still exact."""
    g = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    Q, N, d, kp, n_splits = 5, 300, 24, 16, 3
    keys = torch.randn(N, d, generator=g) * (0.5 + torch.rand(N, 1, generator=g))     # dot-product ranking: norms vary
    q = torch.randn(Q, d, generator=g)
    scale = 1.0 / float(keys.norm(dim=1).max())
    qn = torch.nn.functional.normalize(q, dim=-1)
    ks = keys * scale                                                                  # row norms <= 1
    approx = (qn.half().double() @ ks.half().double().T).numpy()
    exact = (qn.double() @ ks.double().T).numpy()
    eps = 2 * 2.0 ** -11 + 2.0 ** -15
    assert np.abs(approx - exact).max() <= eps
    bounds = np.linspace(0, N, n_splits + 1).astype(int)
    n_cert = 0
    for r in range(Q):
        n_ex = int(rng.integers(0, 40))
        excl = set(np.argsort(-exact[r])[:n_ex].tolist()) if heavy else set(rng.choice(N, n_ex, replace=False).tolist())
        cand, tmax = [], -np.inf
        for s in range(n_splits):
            lo, hi = bounds[s], bounds[s + 1]
            order = lo + np.argsort(-approx[r, lo:hi], kind="stable")[:kp]
            cand.extend(order.tolist())
            if len(order) == kp:
                tmax = max(tmax, approx[r, order[-1]])
        adm = np.array([c for c in cand if c not in excl], dtype=np.int64)
        if adm.size < k:
            continue                                                                   # (the product: second pass)
        e = exact[r, adm]
        order = np.lexsort((adm, -e))[:k]
        if not e[order[-1]] > tmax + eps:
            continue
        n_cert += 1
        masked = exact[r].copy()
        masked[list(excl)] = -np.inf
        ok, bad = O.topk_sets_match(adm[order][None, :], masked[None, :], k)
        assert ok, bad
    if not heavy:
        assert n_cert > 0
