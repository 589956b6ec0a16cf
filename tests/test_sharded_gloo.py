"""world_size-2 (and 3) gloo test of the sharded-retrieval host logic on CPU.

The per-rank compute (local top-k, merge, owner gather) is injected from the ORACLE here -- the product
defaults are the CUDA ops; this test covers partitioning, idx offsets, the all-gather exchange, padding of
tiny shards, and owner selection of gathered rows."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ragraph_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Store:
    retrieve_num = 4

    def __init__(self, keys, values, labels):
        self.resource_keys, self.resource_values, self.resource_labels = keys, values, labels
        self.shard_lo = 0


def _local_topk(store, q, k):
    s, i = O.topk(O.cosine_similarity(q, store.resource_keys), k)
    return s, i + store.shard_lo


def _merge(scores, idx, k):
    return O.merge_topk(scores, idx, k)


def _gather_owned(table_local, idx, lo, n_global, out):
    flat, o = idx.reshape(-1), out.reshape(idx.numel(), *table_local.shape[1:])
    mine = (flat >= lo) & (flat < lo + table_local.shape[0])
    o[mine] = table_local[flat[mine] - lo]


def _worker(rank, world, port, N, results):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ragraph_b200.sharded import ShardedRetriever, shard_bounds
        g = torch.Generator().manual_seed(7)
        q = torch.randn(11, 16, generator=g)
        keys = torch.randn(N, 16, generator=g)
        keys[N // 2] = keys[0]                                    # a tie across shards
        values = torch.randn(N, 16, generator=g)
        labels = torch.nn.functional.one_hot(torch.randint(0, 3, (N,), generator=g), 3)   # int64
        lo, hi = shard_bounds(N, world, rank)
        store = _Store(keys[lo:hi], values[lo:hi], labels[lo:hi])
        sr = ShardedRetriever(store, N, local_topk=_local_topk, merge=_merge, gather_owned=_gather_owned)
        emb, lab, scores, idx = sr.retrieve(q, 4)
        ref_s, ref_i, ref_e, ref_l = O.retrieve(q, keys, values, labels, 4)
        ok, bad = O.topk_sets_match(idx.numpy(), O.cosine_similarity_f64(q, keys), 4)
        assert ok, bad
        assert torch.allclose(scores, ref_s, atol=1e-6)
        assert torch.equal(emb, values[idx]) and torch.equal(lab, labels[idx])      # bit exact gathers
        gathered = [None] * world
        dist.all_gather_object(gathered, idx.tolist())
        assert all(g_ == gathered[0] for g_ in gathered)          # every rank holds the same answer
        results[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,N", [(2, 101), (3, 50), (2, 5)])
def test_sharded_retrieve_gloo(world, N):
    port = _free_port()
    with mp.Manager() as m:
        results = m.dict()
        mp.spawn(_worker, args=(world, port, N, results), nprocs=world, join=True)
        assert len(results) == world and all(results.values())
