#!/usr/bin/env python
"""Throughput of a final() weighting (DateBM25F) on the config-2 workload (development tool).
Dates are synthetic: 2000 session dates shared by the documents, one document in ten undated."""
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
batch = s.pack(qs.queries)
rng = np.random.default_rng(1)
add = 5.0e9 + rng.integers(0, 2000, ix.n_docs_all) * 86400.0 + 1.0
add[rng.random(ix.n_docs_all) < 0.1] = np.nan
for _ in range(3):
    eng.search_batch(batch, 10)
t0 = time.perf_counter()
for _ in range(10):
    eng.search_batch(batch, 10)
t_plain = (time.perf_counter() - t0) / 10
eng.set_final_date(add)
for _ in range(2):
    out = eng.search_batch_final(batch, 10)
eng.reset_stats()
t0 = time.perf_counter()
for _ in range(5):
    out = eng.search_batch_final(batch, 10)
t_final = (time.perf_counter() - t0) / 5
st = eng.stats()
print("plain BM25F: %.3f ms per 10k-query batch (%.2f M q/s); date final: %.3f ms (%.2f M q/s); kernels %s"
      % (t_plain * 1e3, 1e-2 / t_plain, t_final * 1e3, 1e-2 / t_final, {k: st[k] for k in ("ms_score", "ms_stream", "n_executes", "postings_stream", "postings_lookup")}))
final, docids, counts, totals = out
# sanity: dated documents first, values descending
ok = all(np.all(np.diff(final[i, :counts[i]]) <= 0) for i in range(0, 10000, 97))
print("descending:", ok, "first row:", final[0, :3], docids[0, :3], int(totals[0]))
eng.set_final_date(None)
