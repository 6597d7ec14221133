"""N > 1 path on CPU: world_size-2 ``gloo`` run of the exchange step (SURVEY.md §8 e).

Each rank holds one document shard with local docids, scores it with corpus-wide statistics (here
with the oracle, since there is no GPU), encodes its local top-k as the engine's 64-bit W11 keys over
GLOBAL docnums, and the ranks exchange with all_gather / all_reduce exactly as ShardedSearcher does;
the merged list must equal the whole-corpus result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from document_search_engine_b200.corpus import make_corpus, make_queries
from document_search_engine_b200.distributed import decode_keys_host, merge_final_host, merge_gathered_host, merge_keys_host
from document_search_engine_b200.searching import make_keys
from oracle.numpy_oracle import NumpyOracle

K = 10


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _local_keys(full_ix, shard, queries):
    """Local top-k keys [Q, K] (global docnums) and local totals [Q] of one shard."""
    o = NumpyOracle(full_ix, shards=[shard])
    keys = np.zeros((len(queries), K), dtype=np.uint64)
    totals = np.zeros(len(queries), dtype=np.int64)
    for i, q in enumerate(queries):
        top, total = o.search(q, limit=K)
        totals[i] = total
        if top:
            s = np.array([t[0] for t in top], dtype=np.float32)
            d = np.array([t[1] for t in top], dtype=np.uint64)
            keys[i, :len(top)] = make_keys(s, d)
    return keys, totals


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ix = make_corpus(4000, 3000, 7, device="cpu")
        qs = make_queries(80, 3000, 8, 1, 4, "mixed", skip_top=0).queries
        shard = ix.shard(rank, world)
        assert shard.doc_base == ix.n_docs_all * rank // world
        keys, totals = _local_keys(ix, shard, qs)
        local = torch.from_numpy(keys.view(np.int64).reshape(-1))
        gathered = torch.empty(world * local.numel(), dtype=torch.int64)
        dist.all_gather_into_tensor(gathered, local)
        tot = torch.from_numpy(totals.copy())
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        merged = merge_keys_host(gathered.numpy().view(np.uint64).reshape(world, len(qs), K), K)
        # the same exchange with ONE collective, as ShardedSearcher.run_plan does it: keys and match counts travel in
        # one span (here with the padding the library's workspace puts between them)
        tot_off = len(qs) * K + 32
        span = np.zeros(tot_off + len(qs), dtype=np.uint64)
        span[:len(qs) * K] = keys.reshape(-1)
        span[tot_off:] = totals.astype(np.uint64)
        g2 = torch.empty(world * span.size, dtype=torch.int64)
        dist.all_gather_into_tensor(g2, torch.from_numpy(span.view(np.int64)))
        merged2, tot2 = merge_gathered_host(g2.numpy().view(np.uint64).reshape(world, span.size), len(qs), K, tot_off)
        assert np.array_equal(merged2, merged) and np.array_equal(tot2.astype(np.int64), tot.numpy())
        if rank == 0:
            scores, docids, counts = decode_keys_host(merged)
            np.savez(out, scores=scores, docids=docids, counts=counts, totals=tot.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_equals_whole(tmp_path):
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    ix = make_corpus(4000, 3000, 7, device="cpu")
    qs = make_queries(80, 3000, 8, 1, 4, "mixed", skip_top=0).queries
    o = NumpyOracle(ix)
    for i, q in enumerate(qs):
        top, total = o.search(q, limit=K)
        assert int(got["totals"][i]) == total
        n = int(got["counts"][i])
        assert n == len(top)
        assert got["docids"][i, :n].tolist() == [d for _, d in top]
        assert got["scores"][i, :n].tolist() == pytest.approx([s for s, _ in top], rel=1e-6)


def test_shards_partition_the_corpus():
    ix = make_corpus(1000, 500, 3, device="cpu")
    for world in (2, 3, 8):
        shards = [ix.shard(g, world) for g in range(world)]
        assert sum(s.n_docs_all for s in shards) == ix.n_docs_all
        assert sum(s.n_postings for s in shards) == ix.n_postings
        assert [s.doc_base for s in shards] == [ix.n_docs_all * g // world for g in range(world)]
        # W8: statistics stay corpus-wide, docids become local
        for s in shards:
            assert s.docids.max(initial=0) < max(1, s.n_docs_all)


def test_key_roundtrip_and_order():
    s = np.array([3.5, 3.5, 1.0, 0.25, 7.0], dtype=np.float32)
    d = np.array([9, 2, 5, 5, 100], dtype=np.uint64)
    keys = make_keys(s, d)
    order = np.argsort(keys)[::-1]
    assert order.tolist() == [4, 1, 0, 2, 3]          # score desc, docnum asc (W11)
    sc, dc, cnt = decode_keys_host(keys[None, :])
    assert sc[0].tolist() == s.tolist() and dc[0].tolist() == d.tolist() and cnt[0] == 5


# ---- the same exchange under a date-ordered weighting: (final value, docnum) lists instead of keys ------------

def _dated(ix):
    from tests.test_date_final import dated_corpus
    return dated_corpus(ix, seed=11)


def _worker_final(rank, world, port, out):
    from document_search_engine_b200 import DescDateBM25F
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ix = _dated(make_corpus(4000, 3000, 7, device="cpu"))
        qs = make_queries(60, 3000, 8, 1, 4, "mixed", skip_top=0).queries
        shard = ix.shard(rank, world)
        # what a rank's device produces: its shard's k best by final value, global docnums, local totals
        o = NumpyOracle(ix, shards=[shard], final_add=DescDateBM25F().doc_final_terms(ix))
        vals = np.full((len(qs), K), -np.inf)
        docs = np.full((len(qs), K), 0xFFFFFFFF, dtype=np.uint32)
        totals = np.zeros(len(qs), dtype=np.int64)
        for i, q in enumerate(qs):
            top, totals[i] = o.search(q, limit=K)
            vals[i, :len(top)] = [t[0] for t in top]
            docs[i, :len(top)] = [t[1] for t in top]
        g_vals = torch.empty(world * vals.size, dtype=torch.float64)
        g_docs = torch.empty(world * docs.size, dtype=torch.int32)
        dist.all_gather_into_tensor(g_vals, torch.from_numpy(vals.reshape(-1)))
        dist.all_gather_into_tensor(g_docs, torch.from_numpy(docs.view(np.int32).reshape(-1)))
        tot = torch.from_numpy(totals.copy())
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        mv, md, mc = merge_final_host(g_vals.numpy().reshape(world, len(qs), K),
                                      g_docs.numpy().view(np.uint32).reshape(world, len(qs), K), K)
        if rank == 0:
            np.savez(out, vals=mv, docids=md, counts=mc, totals=tot.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_final_exchange_equals_whole(tmp_path):
    from document_search_engine_b200 import DescDateBM25F
    out = str(tmp_path / "merged_final.npz")
    mp.spawn(_worker_final, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    ix = _dated(make_corpus(4000, 3000, 7, device="cpu"))
    qs = make_queries(60, 3000, 8, 1, 4, "mixed", skip_top=0).queries
    o = NumpyOracle(ix, final_add=DescDateBM25F().doc_final_terms(ix))
    for i, q in enumerate(qs):
        top, total = o.search(q, limit=K)
        assert int(got["totals"][i]) == total
        n = int(got["counts"][i])
        assert n == len(top)
        assert got["docids"][i, :n].tolist() == [d for _, d in top]
        assert got["vals"][i, :n].tolist() == [v for v, _ in top]
