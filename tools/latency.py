#!/usr/bin/env python
"""Single-query latency through the Searcher facade (the reference's interactive use), config-2 corpus."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from document_search_engine_b200 import BM25F
from document_search_engine_b200.corpus import config_corpus, config_queries
ix = config_corpus(2)
qs = config_queries(2, 400).queries
with ix.searcher(weighting=BM25F) as s:
    for q in qs[:20]:
        s.search(q, limit=10)
    lat = []
    for q in qs:
        t0 = time.perf_counter()
        r = s.search(q, limit=10)
        lat.append((time.perf_counter() - t0) * 1e6)
    lat = np.array(lat)
    print("Searcher.search(q, limit=10): median %.0f us, p90 %.0f us, p99 %.0f us, max %.0f us over %d queries"
          % (np.median(lat), np.percentile(lat, 90), np.percentile(lat, 99), lat.max(), lat.size))
    t0 = time.perf_counter(); page = s.search_page(qs[0], 2, 10); t1 = time.perf_counter()
    print("search_page(q, 2, 10): %.0f us, total %d" % ((t1 - t0) * 1e6, page.total))
