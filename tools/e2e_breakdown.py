#!/usr/bin/env python
"""Host-side breakdown of one bm25f_search_batch call on config 2 (development tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from document_search_engine_b200.corpus import config_corpus, config_queries
from document_search_engine_b200.scoring import BM25F
from document_search_engine_b200.searching import Searcher

ix = config_corpus(2)
qs = config_queries(2, 10000)
s = Searcher(ix, weighting=BM25F)
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
