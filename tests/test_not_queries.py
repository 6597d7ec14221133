"""NOT clauses (reference UI help search-form.html:20-40): ``a NOT b`` parses to
``And([a, Not(b)])``; Whoosh's compound matcher turns the Not children into the excluded side of an
AndNotMatcher.  CPU: the two oracles against each other and a hand-worked case.  GPU: every kernel that
serves NOT (stream, isect; forced CTA/team variants re-route) against the oracle."""
import numpy as np
import pytest

from document_search_engine_b200 import And, BM25F, FlatIndex, Not, Or, QueryParser, Term
from document_search_engine_b200.corpus import config_corpus, make_queries
from document_search_engine_b200.query import GROUP_NOT, NullQuery, UnsupportedQuery, lower
from oracle.numpy_oracle import NumpyOracle
from oracle.whoosh_port import OracleSearcher
from tests.parity import assert_batch_parity, assert_query_parity
from tests.test_oracle_kat import kat1_index


def T(t):
    return Term("f", t)


def test_parser_and_lowering():
    qp = QueryParser("f")
    assert qp.parse("a NOT b") == And([T("a"), Not(T("b"))])
    assert qp.parse("a NOT b c OR d") == And([T("a"), Not(T("b")), Or([T("c"), T("d")])])
    assert str(qp.parse("a NOT b")) == "(f:a AND NOT f:b)"
    leaves, g, kind = lower(And([T("a"), Not(Or([T("b"), T("c")])), Or([T("d"), T("e")])]))
    assert kind == "groups" and g == 2
    assert [(l.text, l.group) for l in leaves] == [("a", 0), ("d", 1), ("e", 1), ("b", GROUP_NOT), ("c", GROUP_NOT)]
    leaves, g, kind = lower(Or([T("a"), T("b"), Not(T("c"))]))
    assert g == 1 and [(l.text, l.group) for l in leaves] == [("a", 0), ("b", 0), ("c", GROUP_NOT)]
    assert lower(And([Not(T("a")), Not(T("b"))]))[2] == "null"        # nothing positive: no hits (Whoosh)
    assert lower(And([T("a"), Not(NullQuery)]))[0][0].text == "a"      # NOT nothing drops out
    with pytest.raises(UnsupportedQuery):
        lower(Not(T("a")))
    with pytest.raises(UnsupportedQuery):
        lower(And([T("a"), Not(And([T("b"), T("c")]))]))


def test_kat1_not_oracles():
    """kat1: a -> docs 0, 2, 3; b -> docs 2, 3 (+ others); hand-checked exclusions."""
    ix = kat1_index()
    o, w = NumpyOracle(ix), OracleSearcher(ix)
    top_a, _ = o.search(T("a"), limit=10)
    docs_b = {d for _, d in o.search(T("b"), limit=10)[0]}
    want = [(s, d) for s, d in top_a if d not in docs_b]
    for orc in (o, w):
        top, total = orc.search(And([T("a"), Not(T("b"))]), limit=10)
        assert total == len(want) and [d for _, d in top] == [d for _, d in want]
        assert [s for s, _ in top] == pytest.approx([s for s, _ in want], rel=1e-12)
        assert orc.search(And([T("a"), Not(T("a"))]), limit=10) == ([], 0)
        top2, total2 = orc.search(And([T("a"), Not(T("nope"))]), limit=10)
        assert total2 == len(top_a) and [d for _, d in top2] == [d for _, d in top_a]


def not_queries(n, n_terms, seed):
    """AND / OR / mixed queries with one or two NOT terms appended (some dense, some unknown)."""
    rng = np.random.default_rng(seed)
    base = make_queries(n, n_terms, seed, 2, 4, "mixed", skip_top=0).queries
    out = []
    for i, q in enumerate(base):
        subs = list(q.subqueries) if isinstance(q, (And, Or)) else [q]
        neg = [Term("body", int(rng.integers(0, 400 if i % 3 else n_terms))) for _ in range(1 + i % 2)]
        nots = [Not(neg[0])] if len(neg) == 1 else [Not(Or(neg))]
        if i % 7 == 0:
            nots.append(Not(Term("body", "no-such-term")))
        cls = type(q) if isinstance(q, (And, Or)) else And
        out.append(cls(subs + nots))
    return out


@pytest.fixture(scope="module")
def cfg1():
    ix = config_corpus(1, device="cpu")
    return ix, NumpyOracle(ix)


def test_oracles_agree_on_not(cfg1):
    ix, o = cfg1
    w = OracleSearcher(ix)
    qs = not_queries(40, 50_000, 99)
    # the generator must produce real exclusions, not only no-ops
    changed = 0
    for q in qs:
        top, total = o.search(q, limit=10)
        top_w, total_w = w.search(q, limit=10)
        assert total == total_w and [d for _, d in top] == [d for _, d in top_w]
        assert [s for s, _ in top] == pytest.approx([s for s, _ in top_w], rel=1e-12)
        pos = type(q)([s for s in q.subqueries if not isinstance(s, Not)])
        changed += o.search(pos, limit=10)[1] != total
    assert changed >= 10


@pytest.mark.gpu
def test_kat1_not_gpu():
    ix = kat1_index()
    o = NumpyOracle(ix)
    with ix.searcher(weighting=BM25F) as s:
        for q in (And([T("a"), Not(T("b"))]), And([T("b"), Not(T("a"))]), And([T("a"), Not(T("a"))]),
                  And([T("a"), Not(T("nope"))]), Or([T("a"), T("b"), Not(T("a"))]),
                  And([T("a"), T("b"), Not(Or([T("c"), T("nope")]))])):
            r = s.search(q, limit=10)
            assert_query_parity(o, q, r.top_n, len(r), 10)
        assert s.search(And([Not(T("a")), Not(T("b"))])).is_empty()


@pytest.mark.gpu
@pytest.mark.parametrize("variant,kw", [(0, {}), (3, {}), (3, {"subtile_docs": 128, "warp_split": 2048}), (5, {}),
                                        (5, {"isect_split": 64}), (4, {}), (1, {}), (0, {"isect_ratio": 1000})])
def test_not_all_routes(cfg1, variant, kw):
    ix, o = cfg1
    qs = not_queries(300, 50_000, 4242)
    with ix.searcher(weighting=BM25F, variant=variant, **kw) as s:
        res = s.search_batch(qs, limit=10)
    assert_batch_parity(o, qs, res, 10)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [1, 32, 100, 150])
def test_not_limits(cfg1, k):
    ix, o = cfg1
    qs = not_queries(100, 50_000, 17)
    with ix.searcher(weighting=BM25F) as s:
        res = s.search_batch(qs, limit=k)
    assert_batch_parity(o, qs, res, k)


@pytest.mark.gpu
def test_not_refused_where_not_served(cfg1):
    """k > 128 and paging run on the CTA kernels, which do not know NOT: a loud error, no silent wrong answer."""
    ix, _ = cfg1
    with ix.searcher(weighting=BM25F) as s:
        with pytest.raises(Exception, match="NOT clauses"):
            s.search_batch([And([Term("body", 5), Not(Term("body", 6))])], limit=500)
