"""Document-sharded search across the GPUs of one box (SURVEY.md §8 e).

One process per GPU (``torch.distributed``, backend ``nccl``).  Rank ``g`` holds the
contiguous document range ``[g*N/G, (g+1)*N/G)`` as its own CSR with local docids and
``doc_base`` set to the range start — the exact analogue of a Whoosh segment with a doc
offset (W8).  Replicated on every rank: the term dictionary, the corpus-wide ``df`` /
``doc_count_all`` / field totals (so idf and avgfl are global) and the query batch.

Per batch the path has ONE exchange step: every rank scores its shard and produces, per
query, a local top-k as 64-bit W11 keys over *global* docnums and a local match count;
then ONE ``all_gather`` of the span of the plan's workspace that holds both the ``Q*k`` keys and the ``Q``
totals, and ``bm25f_merge_gathered`` (merge kernel + a decode kernel that also adds up the totals) selects the
top-k of the ``G*k`` candidates per query on every rank.  PyTorch supplies the NCCL plumbing and the receive buffers; scoring, top-k and
the merge are the library's own kernels.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _ffi
from .searching import Searcher


class _RawCuda:
    """Zero-copy view of library-owned device memory for ``torch.as_tensor``."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False),
                                         "version": 2, "strides": None}


_VIEWS = {}


def device_view(ptr: int, n: int, device: int, typestr: str = "<i8") -> torch.Tensor:
    """Tensor aliasing ``n`` elements at device pointer ``ptr`` (int64 words by default).  Views are cached:
    the library's result buffers live in two grow-only workspaces, so the same few pointers recur."""
    key = (ptr, n, device, typestr)
    v = _VIEWS.get(key)
    if v is None:
        if len(_VIEWS) > 64:
            _VIEWS.clear()
        v = torch.as_tensor(_RawCuda(ptr, n, typestr), device="cuda:%d" % device)
        _VIEWS[key] = v
    return v


def merge_keys_host(gathered: np.ndarray, k: int) -> np.ndarray:
    """Host restatement of the merge kernel for CPU (gloo) tests: ``gathered`` is
    ``[G, Q, k]`` uint64 keys (0 = empty); returns ``[Q, k]`` with the k largest per query."""
    G, Q, _ = gathered.shape
    allk = np.transpose(gathered, (1, 0, 2)).reshape(Q, G * k)
    allk = np.sort(allk.astype(np.uint64), axis=1)[:, ::-1]      # descending = W11 order
    return np.ascontiguousarray(allk[:, :k])


def merge_gathered_host(gathered: np.ndarray, n_queries: int, k: int, totals_offset: int):
    """Host restatement of ``bm25f_merge_gathered`` for CPU (gloo) tests: ``gathered`` is ``[G, span]`` uint64, every
    row a shard's ``[Q * k]`` keys followed (at word ``totals_offset``) by its ``[Q]`` match counts.  Returns
    ``(merged keys [Q, k], totals [Q])``."""
    G = gathered.shape[0]
    keys = gathered[:, :n_queries * k].reshape(G, n_queries, k)
    totals = gathered[:, totals_offset:totals_offset + n_queries].astype(np.uint64).sum(axis=0)
    return merge_keys_host(keys, k), totals


def merge_final_host(vals: np.ndarray, docids: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Host restatement of ``bm25f_merge_final_lists`` for CPU (gloo) tests: ``vals`` / ``docids`` are
    ``[G, Q, k]`` (float64 final values, uint32 global docnums, 0xFFFFFFFF = empty slot); returns the k best
    per query by (value descending, docnum ascending) as ``(vals [Q, k], docids [Q, k], counts [Q])``."""
    G, Q, _ = vals.shape
    v = np.transpose(vals, (1, 0, 2)).reshape(Q, G * k)
    d = np.transpose(docids, (1, 0, 2)).reshape(Q, G * k).astype(np.int64)
    out_v = np.full((Q, k), -np.inf)
    out_d = np.full((Q, k), 0xFFFFFFFF, dtype=np.uint32)
    counts = np.zeros(Q, dtype=np.uint32)
    for q in range(Q):
        ok = d[q] != 0xFFFFFFFF
        order = np.lexsort((d[q][ok], -v[q][ok]))[:k]
        n = order.size
        out_v[q, :n] = v[q][ok][order]
        out_d[q, :n] = d[q][ok][order]
        counts[q] = n
    return out_v, out_d, counts


def decode_keys_host(keys: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """``(scores f32, docids u32, counts)`` from uint64 keys (host mirror of ``bm25f_decode_keys``)."""
    u = (keys >> np.uint64(32)).astype(np.uint32)
    pos = (u & np.uint32(0x80000000)) != 0
    bits = np.where(pos, u & np.uint32(0x7FFFFFFF), ~u)
    scores = bits.astype(np.uint32).view(np.float32).copy()
    docids = (np.uint64(0xFFFFFFFF) - (keys & np.uint64(0xFFFFFFFF))).astype(np.uint32)
    valid = keys != 0
    scores[~valid] = -np.inf
    docids[~valid] = 0xFFFFFFFF
    return scores, docids, valid.sum(axis=-1).astype(np.uint32)


class ShardedSearcher:
    """Rank-local searcher over one document shard plus the cross-GPU merge."""

    def __init__(self, full_ix, rank: Optional[int] = None, world: Optional[int] = None, device: Optional[int] = None,
                 weighting=None, group=None, shard_ix=None, **engine_opts):
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.device = torch.cuda.current_device() if device is None else device
        self.group = group
        if shard_ix is None:
            shard_ix = full_ix if self.world == 1 else full_ix.shard(self.rank, self.world)
        self.shard_ix = shard_ix
        # corpus-wide statistics come from ``full_ix`` (only its df / totals / dictionary are used)
        self.local = Searcher(self.shard_ix, weighting=weighting, device=self.device, stats_ix=full_ix, **engine_opts)
        self.engine = self.local.engine
        self.engine.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        self._bufs = {}

    def pack(self, queries, after_keys=None):
        return self.local.pack(queries, after_keys)

    def _buffers(self, Q: int, k: int, slot: int = 0):
        key = (Q, k, slot)
        b = self._bufs.get(key)
        if b is None:
            dev = "cuda:%d" % self.device
            # the decoded results sit in ONE buffer (totals | scores | docids | counts) so that a step brings them to
            # the host with one copy
            out = torch.empty(self._out_words(Q, k), dtype=torch.int64, device=dev)
            o32 = out.view(torch.int32)
            b = dict(gathered=None,                   # sized on first use from the plan's gather span
                     merged=torch.empty(Q * k, dtype=torch.int64, device=dev),
                     out=out, totals=out[:Q],
                     scores=o32[2 * Q:2 * Q + Q * k].view(torch.float32),
                     docids=o32[2 * Q + Q * k:2 * Q + 2 * Q * k],
                     counts=o32[2 * Q + 2 * Q * k:3 * Q + 2 * Q * k])
            self._bufs[key] = b
        return b

    @staticmethod
    def _out_words(Q: int, k: int) -> int:
        return Q + (2 * Q * k + Q + 1) // 2

    @staticmethod
    def _split_out(h: np.ndarray, Q: int, k: int):
        """``(scores, docids, counts, totals)`` views of one result buffer (int64 words, layout of ``_buffers``)."""
        w32 = h.view(np.uint32)
        return (w32[2 * Q:2 * Q + Q * k].view(np.float32).reshape(Q, k), w32[2 * Q + Q * k:2 * Q + 2 * Q * k].reshape(Q, k),
                w32[2 * Q + 2 * Q * k:3 * Q + 2 * Q * k], h[:Q].view(np.uint64))

    def run_plan(self, plan: _ffi.Plan, slot: int = 0):
        """Execute a prepared plan on this shard, exchange, merge.  Everything is enqueued on
        torch's current stream; returns the device buffers (merged keys, totals, decoded).  ``slot``
        picks one of the buffer sets (two batches can be in flight, see ``search_packed_stream``)."""
        Q, k = plan.n_queries, plan.k
        b = self._buffers(Q, k, slot)
        plan.execute()
        if self.world > 1:
            # ONE collective and two kernels: the plan's keys and match counts sit in one span of its workspace
            base, span, tot_off = plan.gather_span()
            g = b["gathered"]
            if g is None or g.numel() != self.world * span:
                g = b["gathered"] = torch.empty(self.world * span, dtype=torch.int64, device="cuda:%d" % self.device)
            dist.all_gather_into_tensor(g, device_view(base, span, self.device), group=self.group)
            self.engine.merge_gathered(g.data_ptr(), self.world, span, tot_off, Q, k, b["merged"].data_ptr(), b["scores"].data_ptr(),
                                       b["docids"].data_ptr(), b["counts"].data_ptr(), b["totals"].data_ptr())
        else:
            d_keys, d_totals = plan.device_results()
            b["merged"].copy_(device_view(d_keys, Q * k, self.device))
            b["totals"].copy_(device_view(d_totals, Q, self.device))
            self.engine.decode_keys(b["merged"].data_ptr(), Q, k, b["scores"].data_ptr(), b["docids"].data_ptr(),
                                    b["counts"].data_ptr())
        return b

    def _search_packed_final(self, batch: _ffi.PackedBatch, k: int):
        """``search_packed`` under a final() weighting (DateBM25F): the shards exchange their (final value,
        docnum) lists instead of 64-bit keys and ``bm25f_merge_final_lists`` picks the k best."""
        if self.world == 1:
            return self.local.search_packed(batch, k)
        Q = batch.n_queries
        dev = "cuda:%d" % self.device
        key = ("final", Q, k)
        b = self._bufs.get(key)
        if b is None:
            b = dict(vals=torch.empty(self.world * Q * k, dtype=torch.float64, device=dev),
                     docs=torch.empty(self.world * Q * k, dtype=torch.int32, device=dev),
                     totals=torch.empty(Q, dtype=torch.int64, device=dev),
                     out_vals=torch.empty(Q * k, dtype=torch.float64, device=dev),
                     out_docs=torch.empty(Q * k, dtype=torch.int32, device=dev),
                     out_counts=torch.empty(Q, dtype=torch.int32, device=dev))
            self._bufs[key] = b
        plan = self.engine.prepare(batch, k, arena=True)
        try:
            plan.execute()
            d_final, d_docids, d_totals = plan.device_final()
            dist.all_gather_into_tensor(b["vals"], device_view(d_final, Q * k, self.device, "<f8"), group=self.group)
            dist.all_gather_into_tensor(b["docs"], device_view(d_docids, Q * k, self.device, "<i4"), group=self.group)
            b["totals"].copy_(device_view(d_totals, Q, self.device))
            dist.all_reduce(b["totals"], op=dist.ReduceOp.SUM, group=self.group)
            self.engine.merge_final_lists(b["vals"].data_ptr(), b["docs"].data_ptr(), self.world, Q, k,
                                          b["out_vals"].data_ptr(), b["out_docs"].data_ptr(), b["out_counts"].data_ptr())
            final = b["out_vals"].cpu().numpy().reshape(Q, k)
            docids = b["out_docs"].cpu().numpy().view(np.uint32).reshape(Q, k)
            counts = b["out_counts"].cpu().numpy().view(np.uint32)
            totals = b["totals"].cpu().numpy().view(np.uint64)
        finally:
            plan.close()
        return final, docids, counts, totals

    def search_packed(self, batch: _ffi.PackedBatch, k: int):
        """Host buffers in, host buffers out: ``(scores, docids, counts, totals)`` of the whole corpus (float64
        final values instead of scores under a final() weighting)."""
        if self.local.weighting.use_final:
            return self._search_packed_final(batch, k)
        plan = self.engine.prepare(batch, k, arena=True)
        try:
            b = self.run_plan(plan)
            Q = batch.n_queries
            h = self._host_buffers(Q, k)
            h.copy_(b["out"], non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            scores, docids, counts, totals = self._split_out(h.numpy().copy(), Q, k)
        finally:
            plan.close()
        return scores, docids, counts, totals

    def search_packed_stream(self, batches, k: int):
        """``search_packed`` over an iterable of packed batches, pipelined two deep: the host plans and
        uploads batch i + 1 while the GPUs score, exchange and merge batch i.  Yields the result tuples in
        order.  Every rank must iterate the same batches (the collectives pair up in order)."""
        if self.local.weighting.use_final:
            for batch in batches:                        # not pipelined
                yield self._search_packed_final(batch, k)
            return
        if self.world == 1:
            yield from self.local.search_packed_stream(batches, k)
            return
        stream = torch.cuda.current_stream(self.device)
        d2h = self._d2h_stream()          # results travel beside the next batch's kernels, not in front of them
        inflight = None
        slot = 0

        def land(item):
            # (worker thread) wait for the batch's results to reach the pinned buffer and copy them out of it
            plan, h, ev, Q = item
            ev.synchronize()
            return self._split_out(h.numpy().copy(), Q, k)

        # The wait for batch i and the copy of its results (0.9 MB for 10k queries, ~80 us) run on a helper thread
        # while this thread stages and launches batch i + 1: at 8 shards the GPU step is 0.47 ms and the host work
        # of a step was as long.  Both release the GIL (event wait, memcpy).
        pool = self._bufs.get("pool")
        if pool is None:
            from concurrent.futures import ThreadPoolExecutor
            pool = self._bufs["pool"] = ThreadPoolExecutor(max_workers=1, thread_name_prefix="bm25f-results")
        try:
            for batch in batches:
                fut = pool.submit(land, inflight) if inflight is not None else None
                plan = self.engine.prepare(batch, k, arena=True)
                Q = batch.n_queries
                b = self.run_plan(plan, slot)
                h = self._host_buffers(Q, k, slot)
                d2h.wait_stream(stream)
                with torch.cuda.stream(d2h):
                    h.copy_(b["out"], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(d2h)
                cur = (plan, h, ev, Q)
                slot ^= 1
                prev, inflight = inflight, cur
                if fut is not None:
                    try:
                        res = fut.result()
                    finally:
                        prev[0].close()
                    yield res
            if inflight is not None:
                prev, inflight = inflight, None
                try:
                    res = land(prev)
                finally:
                    prev[0].close()
                yield res
        finally:
            if inflight is not None:
                try:
                    inflight[2].synchronize()
                finally:
                    inflight[0].close()

    def _d2h_stream(self):
        s = self._bufs.get("d2h_stream")
        if s is None:
            s = self._bufs["d2h_stream"] = torch.cuda.Stream(self.device)
        return s

    def _host_buffers(self, Q: int, k: int, slot: int = 0):
        key = ("host", Q, k, slot)
        h = self._bufs.get(key)
        if h is None:
            h = self._bufs[key] = torch.empty(self._out_words(Q, k), dtype=torch.int64).pin_memory()
        return h
