"""Batches planned on the device (csrc/plan.cuh) against the same batches planned on the host (option ``host_plan``):
identical results, and the oracle on a sample.  Batches that the device planner does not take (a NOT clause, a
paging bound, more than 8 leaves) must fall through to the host planner inside the same call."""
import numpy as np
import pytest

from document_search_engine_b200 import And, BM25F, Every, Not, NullQuery, Or, Term
from document_search_engine_b200.corpus import make_corpus, make_queries
from document_search_engine_b200.searching import Searcher
from oracle.numpy_oracle import NumpyOracle
from tests.parity import assert_query_parity

pytestmark = pytest.mark.gpu


def run(ix, queries, k, **opts):
    ix._engine_cache.clear()
    s = Searcher(ix, weighting=BM25F, **opts)
    batch = s.pack(queries)
    out = s.engine.search_batch(batch, k)
    st = s.engine.stats()
    # a second batch through the other arena, then the first again: the workspaces are reused
    out2 = s.engine.search_batch(batch, k)
    for a, b in zip(out, out2):
        assert np.array_equal(a, b)
    s.engine.close()
    ix._engine_cache.clear()
    return out, st


def same(a, b):
    sa, da, ca, ta = a
    sb, db, cb, tb = b
    assert np.array_equal(ca, cb)
    assert np.array_equal(ta, tb)
    for i in range(len(ca)):
        n = int(ca[i])
        assert np.array_equal(da[i, :n], db[i, :n]), "query %d docids" % i
        assert np.array_equal(sa[i, :n], sb[i, :n]), "query %d scores" % i


def oracle_sample(ix, queries, out, k, every=37):
    o = NumpyOracle(ix)
    scores, docids, counts, totals = out
    for i in range(0, len(queries), every):
        n = int(counts[i])
        assert_query_parity(o, queries[i], list(zip(scores[i, :n].tolist(), docids[i, :n].tolist())), int(totals[i]), k,
                            ctx="query %d" % i)


@pytest.mark.parametrize("mode,k", [("or", 10), ("and", 10), ("mixed", 10), ("mixed", 100), ("or", 256)])
def test_device_plan_equals_host_plan(mode, k):
    ix = make_corpus(60000, 3000, 11, device="cpu")
    queries = make_queries(700, 3000, 5, 1, 8, mode).queries
    # dead groups, unknown terms, empty queries and Every() inside the batch
    queries[3] = And([Term("body", 5), Term("body", "nope")])
    queries[4] = Or([Term("body", "nope"), Term("nofield", 1)])
    queries[5] = NullQuery
    queries[6] = Every("body")
    queries[7] = And([Or([Term("body", 1), Term("body", 2)]), Or([Term("body", 2), Term("body", 1)])])
    dev, st_dev = run(ix, queries, k)
    host, st_host = run(ix, queries, k, host_plan=1)
    same(dev, host)
    assert st_dev["postings_touched"] == st_host["postings_touched"]
    for name in ("postings_stream", "postings_team", "postings_lookup"):
        assert st_dev[name] == st_host[name], name
    assert st_dev["n_launches"] >= st_host["n_launches"] + 3          # the three plan kernels are counted
    oracle_sample(ix, queries, dev, k)


def test_heavy_queries_are_cut_the_same_way():
    """Long posting lists: every query is cut into several items by both planners."""
    ix = make_corpus(300000, 120, 13, device="cpu")       # 120 terms: every list is long
    queries = make_queries(300, 120, 9, 4, 8, "or", skip_top=0).queries
    dev, st_dev = run(ix, queries, 10)
    host, st_host = run(ix, queries, 10, host_plan=1)
    same(dev, host)
    assert st_dev["n_items"] > len(queries)
    oracle_sample(ix, queries, dev, 10, every=59)


def test_ineligible_batches_fall_through_to_the_host_planner():
    ix = make_corpus(20000, 800, 17, device="cpu")
    base = make_queries(400, 800, 7, 1, 6, "mixed").queries
    with_not = list(base)
    with_not[10] = And([Term("body", 3), Not(Term("body", 4))])
    wide = list(base)
    wide[11] = Or([Term("body", i) for i in range(12)])
    for queries in (with_not, wide):
        dev, _ = run(ix, queries, 10)
        host, _ = run(ix, queries, 10, host_plan=1)
        same(dev, host)
        oracle_sample(ix, queries, dev, 10, every=10)


def test_submit_collect_with_device_plans():
    ix = make_corpus(40000, 2000, 19, device="cpu")
    s = Searcher(ix, weighting=BM25F)
    batches = [s.pack(make_queries(512, 2000, 100 + i, 1, 8, "mixed").queries) for i in range(4)]
    want = [s.engine.search_batch(b, 10) for b in batches]
    pend = [s.engine.submit(batches[0], 10)]
    got = []
    for i in range(1, 4):
        pend.append(s.engine.submit(batches[i], 10))
        got.append(pend.pop(0).collect())
    got.append(pend.pop(0).collect())
    for a, b in zip(got, want):
        same(a, b)
    s.engine.close()
    ix._engine_cache.clear()
