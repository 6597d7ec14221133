"""Edge cases of the CUDA path through the C ABI: empty and ragged inputs, degenerate indexes, repeated
terms, extreme weights.  Every case is compared with the oracle."""
import numpy as np
import pytest

from document_search_engine_b200 import And, BM25F, Every, FlatIndex, Not, NullQuery, Or, Term
from document_search_engine_b200 import _ffi
from document_search_engine_b200.corpus import make_corpus, make_queries
from oracle.numpy_oracle import NumpyOracle
from tests.parity import assert_batch_parity, assert_query_parity

pytestmark = pytest.mark.gpu
VARIANTS = [0, 3, 4, 5, 1]


def check(ix, queries, limit=10, **kw):
    o = NumpyOracle(ix)
    with ix.searcher(**kw) as s:
        res = s.search_batch(queries, limit=limit)
    assert len(res) == len(queries)
    assert_batch_parity(o, queries, res, limit)
    ix._engine_cache.clear()
    return res


@pytest.mark.parametrize("variant", VARIANTS)
def test_empty_batch_and_null_queries(variant):
    ix = make_corpus(500, 200, 3, device="cpu")
    with ix.searcher(variant=variant) as s:
        assert s.search_batch([], limit=10) == []
        empty = _ffi.PackedBatch([0], [], [], [], [])
        sc, dc, cn, tt = s.engine.search_batch(empty, 10)
        assert sc.shape == (0, 10) and tt.size == 0
    ix._engine_cache.clear()
    qs = [Term("body", "nope"), Or([Term("body", "nope"), Term("nofield", 3)]), And([Term("body", 5), Term("body", "nope")]),
          NullQuery, Or([]), And([]), Term("body", 5)]
    res = check(ix, qs, variant=variant)
    assert [len(r) for r in res[:6]] == [0] * 6 and len(res[6]) > 0


@pytest.mark.parametrize("variant", VARIANTS)
def test_degenerate_indexes(variant):
    one = FlatIndex.from_documents([{"f": ["a", "b", "a"]}], ["f"])
    check(one, [Term("f", "a"), And([Term("f", "a"), Term("f", "b")]), Or([Term("f", "b"), Term("f", "c")]), Every("f")], variant=variant)
    gone = FlatIndex.from_documents([{"f": ["a"]}, {"f": ["a", "b"]}, {"f": ["b"]}], ["f"], deleted=[0, 1, 2])
    res = check(gone, [Term("f", "a"), Or([Term("f", "a"), Term("f", "b")]), Every("f")], variant=variant)
    assert all(r.is_empty() for r in res)
    no_postings = FlatIndex.from_documents([{"f": []}, {"f": []}], ["f"])
    res = check(no_postings, [Term("f", "a"), Every("f")], variant=variant)
    assert all(r.is_empty() for r in res)


@pytest.mark.parametrize("variant", VARIANTS)
def test_repeated_terms_and_extreme_boosts(variant):
    ix = make_corpus(4000, 300, 21, device="cpu")
    qs = [Or([Term("body", 3), Term("body", 3)]),                       # the same list twice: scores add (W10)
          And([Term("body", 3), Term("body", 3), Term("body", 7)]),
          And([Or([Term("body", 3), Term("body", 4)]), Or([Term("body", 4), Term("body", 3)])]),
          Or([Term("body", 2, boost=1e6), Term("body", 250, boost=1e-6)]),
          And([Term("body", 1, boost=1e-20), Term("body", 2)]),           # tiny but positive weight
          Or([Term("body", 299), Term("body", 298), Term("body", 297), Term("body", 296), Term("body", 295),
              Term("body", 294), Term("body", 293), Term("body", 292)]),  # eight rare leaves
          Or([Term("body", i) for i in range(40, 52)])]                   # twelve leaves: past the stream kernels' eight
    check(ix, qs, variant=variant)
    check(ix, qs, limit=1, variant=variant)
    check(ix, qs, limit=128, variant=variant)


def test_ragged_batch_sizes_and_splits():
    """Batches of 1, 2, 1023, 1025 queries (the small-batch item splitting switches at 1024) and a
    dense query cut into hundreds of items."""
    ix = make_corpus(30000, 2000, 8, device="cpu")
    o = NumpyOracle(ix)
    qs = make_queries(1025, 2000, 4, 1, 4, "mixed", skip_top=0).queries
    with ix.searcher() as s:
        for n in (1, 2, 1023, 1025):
            res = s.search_batch(qs[:n], limit=10)
            assert_batch_parity(o, qs[:n], res, 10, sample=range(0, n, max(1, n // 60)))
    ix._engine_cache.clear()
    dense = [Or([Term("body", 0), Term("body", 1), Term("body", 2)]), And([Term("body", 0), Term("body", 1)])]
    with ix.searcher(warp_split=256, isect_split=128, cta_split=256, variant=0) as s:
        res = s.search_batch(dense, limit=10)
        assert s.engine.stats()["n_items"] > 100
    assert_batch_parity(o, dense, res, 10)


def test_bad_arguments_raise():
    ix = make_corpus(200, 100, 1, device="cpu")
    with ix.searcher() as s:
        with pytest.raises(_ffi.EngineError):
            s.engine.search_batch(_ffi.PackedBatch([0, 1], [1], [10 ** 6], [1.0], [0]), 10)      # term id out of range
        with pytest.raises(_ffi.EngineError):
            s.engine.search_batch(_ffi.PackedBatch([0, 2], [2], [1, 2], [1.0, 1.0], [1, 0]), 10)  # groups must not decrease
        with pytest.raises(_ffi.EngineError):
            s.engine.search_batch(_ffi.PackedBatch([0, 1], [1], [1], [float("nan")], [0]), 10)
        with pytest.raises(_ffi.EngineError):
            s.engine.search_batch(_ffi.PackedBatch([0, 1], [1], [1], [1.0], [0]), 0)             # k out of range
        with pytest.raises(ValueError):
            s.search(Term("body", 1), limit=0)
        # the engine is still usable afterwards
        assert len(s.search(Term("body", 1))) == int(ix.df[ix.term_id("body", 1)])
    with pytest.raises(_ffi.EngineError):
        ix._engine_cache.clear()
        ix.searcher(subtile_docs=100)                                                          # not a multiple of 128


@pytest.mark.parametrize("variant", VARIANTS)
def test_non_scorable_field_w15(variant):
    """W15: terms of a field that is not scorable (the reference's ``book=ID``, my_index.py:152; the UI's book filter
    writes ``book:xyz`` / ``NOT book:xyz``, static/main.js:5-16) score their posting weight times the boost
    (Whoosh's WeightScorer): alone, summed with BM25F leaves in And / Or, and as NOT clauses."""
    rng = np.random.default_rng(7)
    books = ["ss", "nopr", "deavf1", "tes%d" % 1, "tes%d" % 2]
    docs = []
    for d in range(3000):
        toks = ["w%d" % t for t in rng.zipf(1.3, size=int(rng.integers(5, 60))) if t < 400]
        docs.append({"body": toks or ["w1"], "book": books[int(rng.integers(0, len(books)))]})
    ix = FlatIndex.from_documents(docs, ["body", "book"], id_fields=["book"], deleted=[5, 77, 1234])
    qs = [Term("book", "nopr"), Term("book", "nopr", boost=2.5), Term("book", "none"),
          Or([Term("book", "ss"), Term("book", "tes1")]),
          And([Term("body", "w1"), Term("book", "ss")]), And([Term("body", "w2"), Term("body", "w3"), Term("book", "nopr", boost=0.25)]),
          Or([Term("body", "w5"), Term("book", "deavf1")]),
          And([Term("body", "w1"), Not(Term("book", "ss"))]), And([Term("body", "w1"), Not(Or([Term("book", "ss"), Term("book", "nopr")]))]),
          And([Or([Term("body", "w2"), Term("body", "w9")]), Or([Term("book", "tes1"), Term("book", "tes2")])])]
    for k in (10, 150):
        check(ix, qs, limit=k, variant=variant)


def test_pattern_queries_f3():
    """Prefix / Wildcard on the device path: the host expands them into Or-of-Terms, the flat-OR / group kernels score them."""
    from tests.test_boundary_cpu import PATTERN_QUERIES, _books_index
    ix = _books_index()
    for variant in (0, 5):
        check(ix, PATTERN_QUERIES, variant=variant)


def test_more_like_f4():
    """searcher.more_like(docnum, field, text=..., top=5) (reference my_flask.py:431-434): key terms -> boosted Or on
    the GPU, the document itself masked out; against the oracle's more_like."""
    from oracle.whoosh_port import OracleSearcher
    from tests.test_boundary_cpu import _books_index
    ix = _books_index()
    o = OracleSearcher(ix)
    with ix.searcher() as s:
        for docnum, text in ((17, None), (3, "w1 w7 w7 w150 w3 w9"), (None, "w2 w2 w5 w11"), (250, "nosuchword")):
            r = s.more_like(docnum, "body", text=text, top=5)
            vec = [(t, 1) for t in text.split()] if text else ix.doc_terms(docnum, "body")
            want_top, want_total = o.more_like(docnum, "body", vec, top=5)
            assert len(r) == want_total and [d for _, d in r.top_n] == [d for _, d in want_top]
            assert [x for x, _ in r.top_n] == pytest.approx([x for x, _ in want_top], rel=1e-5)
            assert docnum not in [h.docnum for h in r]
    ix._engine_cache.clear()


def test_malformed_posting_lists_are_refused():
    """bm25f_create checks what the header only used to document: docids < n_docs_all, strictly ascending inside a list."""
    ix = make_corpus(300, 100, 5, device="cpu")
    for kind in ("range", "order", "dup"):
        bad = make_corpus(300, 100, 5, device="cpu")
        a, b = int(bad.term_offsets[40]), int(bad.term_offsets[41])
        assert b - a >= 3
        if kind == "range":
            bad.docids[b - 1] = 300
        elif kind == "order":
            bad.docids[a], bad.docids[a + 1] = bad.docids[a + 1], bad.docids[a]
        else:
            bad.docids[a + 1] = bad.docids[a]
        with pytest.raises(_ffi.EngineError) as e:
            bad.searcher()
        assert e.value.code == -1 and "posting list 40" in str(e.value)
    with ix.searcher() as s:
        assert len(s.search(Term("body", 3))) > 0
    ix._engine_cache.clear()


def test_compact_store_releases_the_raw_postings():
    """Option compact_store: 8 of the 16 bytes a posting are released after the first set_weighting; the warp kernels
    serve the same results, anything that needs the raw postings is refused loudly."""
    ix = make_corpus(30000, 1500, 23, device="cpu")
    qs = make_queries(300, 1500, 3, 1, 6, "mixed").queries
    qs[5] = And([Term("body", 3), Not(Term("body", 4))])
    with ix.searcher() as s:
        full_bytes = s.engine.stats()["device_bytes"]
    ix._engine_cache.clear()
    o = NumpyOracle(ix)
    with ix.searcher(compact_store=1) as s:
        assert s.engine.stats()["device_bytes"] < 0.62 * full_bytes
        for limit in (10, 100):
            res = s.search_batch(qs, limit=limit)
            assert_batch_parity(o, qs, res, limit)
        with pytest.raises(RuntimeError, match="compact_store"):
            s.search_batch(qs[:4], limit=300)                     # k > 256: the CTA kernels read the raw store
        with pytest.raises(RuntimeError, match="compact_store"):
            ix.searcher(weighting=BM25F(B=0.3), compact_store=1).search(qs[0], limit=10)   # another weighting, same engine
        res = s.search_batch(qs, limit=10)                        # the engine still serves its weighting
        assert_batch_parity(o, qs, res, 10)
    ix._engine_cache.clear()


def test_date_ranges_f3():
    """``date:[oct 1970 to dec 8 1970]`` / ``hospital date:"feb 1964"`` (search-form.html:26, :39): DateRange queries
    expand on the host into the index's year / month / day posting lists and run on the ordinary kernels; the oracle
    evaluates them over the stored dates."""
    from tests.test_dates import corpus, queries
    from document_search_engine_b200 import QueryParser
    ix = corpus(3000, seed=11)
    qp = QueryParser("body")
    qs = queries() + [qp.parse('w5 date:"feb 1966"'), qp.parse("date:[oct 1965 to dec 8 1967]"), qp.parse("w1 OR w2 date:1968"),
                      qp.parse("date:[to 1964]"), qp.parse("w3 NOT date:[1965 to 1970]")]
    for limit in (10, 100):
        res = check(ix, qs, limit=limit)
    assert len(res[0]) > 0 and res[4].is_empty()
    one = check(ix, [queries()[5]], limit=5)[0]                 # a single day, boost 2.5: every hit scores the boost
    assert all(abs(h.score - 2.5) < 1e-6 for h in one)


def test_phrases_f3():
    """Quoted phrases (search-form.html:20-40; Whoosh query.Phrase): the host finds the documents that pass the
    positional test, the library takes them as per-batch posting lists (bm25f_put_lists) and scores
    And(words, list); the oracle restates Whoosh's SpanNear semantics over the stored word order."""
    from tests.test_phrases import corpus, phrases
    from document_search_engine_b200 import Phrase, QueryParser
    ix = corpus(5000, seed=21, vocab=30)
    qp = QueryParser("body")
    qs = phrases() + [And([Term("body", "w0"), Phrase("body", ["w1", "w2"])]),
                      And([Phrase("body", ["w1", "w2"]), Phrase("body", ["w2", "w1"], boost=3.0)]),
                      And([Term("body", "w0"), Not(Phrase("body", ["w1", "w2"]))]),
                      qp.parse('w3 "w4 w5"~2 NOT "w6 w7"'), qp.parse('title:"w1 w2" w3'), Term("body", "w1"),
                      Or([Term("body", "w1"), Term("body", "w2")])]
    for limit in (10, 100):
        res = check(ix, qs, limit=limit)
    assert len(res[0]) > 0 and res[6].is_empty() and res[8].is_empty()
    # many queries in one batch (the device planner's path), every phrase its own list
    many = [Phrase("body", ["w%d" % (i % 30), "w%d" % ((i * 7 + 1) % 30)], slop=1 + i % 3) for i in range(300)]
    check(ix, many, limit=10)
    # pre-packed batches cannot carry per-batch lists
    from document_search_engine_b200.query import UnsupportedQuery
    with ix.searcher() as s:
        with pytest.raises(UnsupportedQuery):
            s.pack([Phrase("body", ["w1", "w2"])])
        page = s.search_page(qp.parse('"w1 w2"'), 2, 5)
        assert page.pagenum == 2 and page.total == len(NumpyOracle(ix).match_all(Phrase("body", ["w1", "w2"]))[0])
    ix._engine_cache.clear()
