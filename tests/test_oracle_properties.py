"""Property tests on random tiny indexes (test tier 1): the doc-at-a-time and the
term-at-a-time oracle agree, AND/OR obey set algebra, document shards merge to the whole (W8),
variant expansion only ever adds matches."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from document_search_engine_b200 import And, FlatIndex, Not, Or, Term
from document_search_engine_b200.variants import expand_with_map
from oracle.numpy_oracle import NumpyOracle
from oracle.whoosh_port import OracleSearcher

WORDS = ["w%d" % i for i in range(12)]
docs_st = st.lists(st.lists(st.sampled_from(WORDS), min_size=0, max_size=12), min_size=1, max_size=24)
term_st = st.builds(lambda w, b: Term("f", w, boost=b), st.sampled_from(WORDS + ["zz"]), st.sampled_from([1.0, 1.0, 2.0, 0.5]))
or_st = st.builds(Or, st.lists(term_st, min_size=1, max_size=3))
# the normal form the engine scores (SURVEY.md §8 b): a term, an OR of terms, or an AND of those
query_st = st.one_of(term_st, or_st, st.builds(And, st.lists(st.one_of(term_st, or_st), min_size=1, max_size=3)))
# ... plus NOT clauses inside AND / OR (Whoosh's AndNotMatcher)
not_st = st.builds(Not, st.one_of(term_st, or_st))
not_query_st = st.one_of(
    st.builds(lambda pos, neg: And(pos + neg), st.lists(st.one_of(term_st, or_st), min_size=1, max_size=3), st.lists(not_st, min_size=1, max_size=2)),
    st.builds(lambda pos, neg: Or(pos + neg), st.lists(term_st, min_size=1, max_size=3), st.lists(not_st, min_size=1, max_size=2)))


def build(docs, deleted_mask):
    ix = FlatIndex.from_documents([{"f": d} for d in docs], ["f"],
                                  deleted=[i for i, x in enumerate(deleted_mask[:len(docs)]) if x])
    return ix


@settings(max_examples=120, deadline=None)
@given(docs_st, st.lists(st.booleans(), min_size=24, max_size=24), query_st, st.sampled_from([1, 3, 10, None]))
def test_daat_and_taat_agree(docs, dele, q, limit):
    ix = build(docs, dele)
    a_top, a_tot = OracleSearcher(ix).search(q, limit=limit)
    b_top, b_tot = NumpyOracle(ix).search(q, limit=limit)
    assert a_tot == b_tot
    assert [d for _, d in a_top] == [d for _, d in b_top]
    assert [s for s, _ in a_top] == pytest.approx([s for s, _ in b_top], rel=1e-12)
    # W11: score descending, docnum ascending among equals
    for (s1, d1), (s2, d2) in zip(a_top, a_top[1:]):
        assert s1 > s2 or (s1 == s2 and d1 < d2)


@settings(max_examples=80, deadline=None)
@given(docs_st, st.lists(st.booleans(), min_size=24, max_size=24), st.sampled_from(WORDS), st.sampled_from(WORDS))
def test_set_algebra(docs, dele, w1, w2):
    ix = build(docs, dele)
    o = NumpyOracle(ix)
    a, b = Term("f", w1), Term("f", w2)
    da, _ = o.match_all(a)
    db, _ = o.match_all(b)
    d_or, s_or = o.match_all(Or([a, b]))
    d_and, s_and = o.match_all(And([a, b]))
    assert set(d_or.tolist()) == set(da.tolist()) | set(db.tolist())
    assert set(d_and.tolist()) == set(da.tolist()) & set(db.tolist())
    assert d_or.size + d_and.size == da.size + db.size
    # a document in both scores the same under AND and OR (W10: sum of the children's scores)
    both = dict(zip(d_and.tolist(), s_and.tolist()))
    for d, s in zip(d_or.tolist(), s_or.tolist()):
        if d in both:
            assert s == pytest.approx(both[d], rel=1e-12)


@settings(max_examples=60, deadline=None)
@given(docs_st, query_st, st.integers(min_value=2, max_value=4))
def test_shards_merge_to_whole(docs, q, n_shards):
    ix = build(docs, [False] * 24)
    whole, total = NumpyOracle(ix).search(q, limit=5)
    parts, tot = [], 0
    for g in range(n_shards):
        top, t = NumpyOracle(ix, shards=[ix.shard(g, n_shards)]).search(q, limit=5)
        parts.extend(top)
        tot += t
    merged = sorted(parts, key=lambda x: (-x[0], x[1]))[:5]
    assert tot == total
    assert [d for _, d in merged] == [d for _, d in whole]
    assert [s for s, _ in merged] == pytest.approx([s for s, _ in whole], rel=1e-12)


@settings(max_examples=60, deadline=None)
@given(docs_st, query_st)
def test_variant_expansion_only_adds_matches(docs, q):
    ix = build(docs, [False] * 24)
    partner = {WORDS[i]: WORDS[i ^ 1] for i in range(len(WORDS))}
    o = OracleSearcher(ix)                      # the doc-at-a-time port takes arbitrary trees (nested ORs)
    t0, n0 = o.search(q, limit=None)
    t1, n1 = o.search(expand_with_map(q, lambda w: partner.get(w)), limit=None)
    assert n0 <= n1 and {d for _, d in t0} <= {d for _, d in t1}


@settings(max_examples=120, deadline=None)
@given(docs_st, st.lists(st.booleans(), min_size=24, max_size=24), not_query_st, st.sampled_from([1, 3, None]))
def test_not_clauses(docs, dele, q, limit):
    """AndNot: the two oracles agree, and the matches are exactly the matches of the positive part that are
    in none of the negated queries, with the positive part's scores."""
    ix = build(docs, dele)
    a_top, a_tot = OracleSearcher(ix).search(q, limit=limit)
    b_top, b_tot = NumpyOracle(ix).search(q, limit=limit)
    assert a_tot == b_tot
    assert [d for _, d in a_top] == [d for _, d in b_top]
    assert [s for s, _ in a_top] == pytest.approx([s for s, _ in b_top], rel=1e-12)
    o = NumpyOracle(ix)
    pos = type(q)([s for s in q.subqueries if not isinstance(s, Not)])
    d_pos, s_pos = o.match_all(pos)
    excluded = set()
    for s in q.subqueries:
        if isinstance(s, Not):
            excluded |= set(o.match_all(s.query)[0].tolist())
    d_q, s_q = o.match_all(q)
    want = [(d, sc) for d, sc in zip(d_pos.tolist(), s_pos.tolist()) if d not in excluded]
    assert d_q.tolist() == [d for d, _ in want]
    assert s_q.tolist() == pytest.approx([sc for _, sc in want], rel=1e-12)
