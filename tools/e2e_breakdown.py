#!/usr/bin/env python
"""Host-side breakdown of one bm25f_search_batch call (development tool): ``e2e_breakdown.py [config] [host_plan]``.
Run it with BM25F_TRACE=1 to see the library's own planner / gap / phase timings per call (first thing to do for the
open config-3 end-to-end defect noted in DESIGN.md section 5)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from document_search_engine_b200.corpus import config_corpus, config_queries
from document_search_engine_b200.scoring import BM25F
from document_search_engine_b200.searching import Searcher

CONFIG = int(sys.argv[1]) if len(sys.argv) > 1 else 2
OPTS = {"host_plan": 1} if len(sys.argv) > 2 and sys.argv[2] == "host_plan" else {}
ix = config_corpus(CONFIG)
qs = config_queries(CONFIG, 10000)
s = Searcher(ix, weighting=BM25F, **OPTS)
eng = s.engine
t0 = time.perf_counter(); batch = s.pack(qs.queries); t1 = time.perf_counter()
print("python pack (query trees -> arrays): %.2f ms" % ((t1 - t0) * 1e3))
for _ in range(3):
    eng.search_batch(batch, 10)
n = 20
t0 = time.perf_counter()
for _ in range(n):
    p = eng.prepare(batch, 10); p.close()
t1 = time.perf_counter()
print("bm25f_prepare (+destroy, own allocations): %.3f ms" % ((t1 - t0) * 1e3 / n))
t0 = time.perf_counter()
for _ in range(n):
    p = eng.prepare(batch, 10, arena=True); eng.synchronize(); p.close()
t1 = time.perf_counter()
print("bm25f_prepare_arena + synchronize (device planner when eligible): %.3f ms" % ((t1 - t0) * 1e3 / n))
t0 = time.perf_counter()
for _ in range(n):
    p = eng.prepare(batch, 10, arena=True); p.execute(); eng.synchronize(); p.close()
t1 = time.perf_counter()
print("prepare_arena + execute + synchronize, a fresh plan every time: %.3f ms; stats %s" % ((t1 - t0) * 1e3 / n, eng.stats()))
p = eng.prepare(batch, 10)
t0 = time.perf_counter()
for _ in range(n):
    p.execute(); eng.synchronize()
t1 = time.perf_counter()
print("execute + synchronize: %.3f ms" % ((t1 - t0) * 1e3 / n))
t0 = time.perf_counter()
for _ in range(n):
    p.fetch()
t1 = time.perf_counter()
print("fetch (D2H + numpy alloc): %.3f ms" % ((t1 - t0) * 1e3 / n))
t0 = time.perf_counter()
for _ in range(n):
    eng.search_batch(batch, 10)
t1 = time.perf_counter()
print("bm25f_search_batch total: %.3f ms" % ((t1 - t0) * 1e3 / n))
