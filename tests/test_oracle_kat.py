"""Oracle vs known-answer vectors (SURVEY.md §8 c, KAT-1 and the listed extras)."""
import math

import numpy as np
import pytest

from document_search_engine_b200 import And, Every, FlatIndex, Or, Term
from document_search_engine_b200.numeric import B2L, length_to_byte, lengths_to_bytes
from oracle.numpy_oracle import NumpyOracle
from oracle.whoosh_port import OracleSearcher, byte_to_length, page_view
from oracle import whoosh_port


def kat1_index(deleted=()):
    # exact lengths [3, 5, 10, 50]; postings a -> {0:1, 2:2, 3:5}, b -> {2:1, 3:1}
    docs = [
        {"f": ["a", "x0", "x1"]},
        {"f": ["y0", "y1", "y2", "y3", "y4"]},
        {"f": ["a", "a", "b"] + ["z%d" % i for i in range(7)]},
        {"f": ["a"] * 5 + ["b"] + ["w%d" % i for i in range(44)]},
    ]
    return FlatIndex.from_documents(docs, ["f"], deleted=deleted)


KAT1_OR = [(2, 3.1036244074953756), (3, 2.1209246397136083), (0, 1.5080645161290325)]
KAT1_AND = [(2, 3.1036244074953756), (3, 2.1209246397136083)]


@pytest.mark.parametrize("oracle_cls", [OracleSearcher, NumpyOracle])
def test_kat1(oracle_cls):
    ix = kat1_index()
    assert ix.len_bytes[0].tolist() == [3, 5, 10, 32]
    assert [int(B2L[b]) for b in ix.len_bytes[0]] == [3, 5, 10, 49]
    assert ix.avg_field_length("f") == 17.0
    o = oracle_cls(ix)
    assert o.idf("f", "a") == 1.0
    assert o.idf("f", "b") == 1.2876820724517808
    top, total = o.search(Or([Term("f", "a"), Term("f", "b")]), limit=10)
    assert total == 3
    assert [d for _, d in top] == [d for d, _ in KAT1_OR]
    for (s, _), (_, want) in zip(top, KAT1_OR):
        assert s == pytest.approx(want, rel=1e-15)
    top, total = o.search(And([Term("f", "a"), Term("f", "b")]), limit=10)
    assert total == 2
    assert [(d, pytest.approx(s, rel=1e-15)) for s, d in top] == KAT1_AND


def test_kat1_leaf_scores():
    o = NumpyOracle(kat1_index())
    d, s = o.leaf_scores(o.ix, "f", "a", 1.0)
    assert d.tolist() == [0, 2, 3]
    assert s.tolist() == pytest.approx([1.5080645161290325, 1.5550935550935554, 1.3934426229508197], rel=1e-15)
    d, s = o.leaf_scores(o.ix, "f", "b", 1.0)
    assert s.tolist() == pytest.approx([1.5485308524018202, 0.7274820167627887], rel=1e-15)


def test_length_tables():
    # W6: constants and probes recorded in SURVEY.md §8 a5
    assert length_to_byte(108116) == 255 and length_to_byte(108115) == 255 or length_to_byte(108115) <= 255
    assert length_to_byte(None) == 0
    for L in range(0, 11):
        assert int(B2L[length_to_byte(L)]) == L
    probes = {20: 20, 50: 49, 100: 101, 180: 182, 500: 508, 1000: 998, 4096: 4112}
    for L, want in probes.items():
        assert int(B2L[length_to_byte(L)]) == want
    assert np.all(np.diff(B2L[1:]) > 0)
    assert [byte_to_length(b) for b in range(256)] == B2L.tolist()
    assert [whoosh_port.length_to_byte(L) for L in range(0, 5000, 7)] == [length_to_byte(L) for L in range(0, 5000, 7)]
    ls = np.array([0, 1, 5, 27, 100, 4096, 108115, 108116, 10 ** 7])
    assert lengths_to_bytes(ls).tolist() == [length_to_byte(int(x)) for x in ls]
    for L in (100, 1000, 50000):
        assert abs(int(B2L[length_to_byte(L)]) - L) / L <= 0.091


@pytest.mark.parametrize("oracle_cls", [OracleSearcher, NumpyOracle])
def test_extras(oracle_cls):
    ix = kat1_index()
    o = oracle_cls(ix)
    a, b = Term("f", "a"), Term("f", "b")
    # unknown term / unknown field: empty matcher, not an error (W10)
    assert o.search(Term("f", "nope")) == ([], 0)
    assert o.search(Term("nofield", "a")) == ([], 0)
    assert o.search(And([a, Term("f", "nope")])) == ([], 0)
    top, total = o.search(Or([a, Term("f", "nope")]))
    assert total == 3
    # leaf boost multiplies the leaf score (W10)
    top2, _ = o.search(Or([Term("f", "a", boost=2.0), b]))
    s = dict((d, v) for v, d in top2)
    assert s[0] == pytest.approx(2 * 1.5080645161290325, rel=1e-15)
    assert s[2] == pytest.approx(2 * 1.5550935550935554 + 1.5485308524018202, rel=1e-15)
    # limit cuts the list but not the total (W11, W13)
    top, total = o.search(Or([a, b]), limit=2)
    assert total == 3 and [d for _, d in top] == [2, 3]
    # limit=None returns every match in the same order (W11)
    top, total = o.search(Or([a, b]), limit=None)
    assert [d for _, d in top] == [2, 3, 0]


@pytest.mark.parametrize("oracle_cls", [OracleSearcher, NumpyOracle])
def test_deleted_doc(oracle_cls):
    # W9: filtered from matches, still counted in dc / df / field length
    ix = kat1_index(deleted=[2])
    assert ix.doc_count_all() == 4 and ix.doc_count() == 3
    o = oracle_cls(ix)
    assert o.idf("f", "b") == 1.2876820724517808
    top, total = o.search(Or([Term("f", "a"), Term("f", "b")]))
    assert total == 2 and [d for _, d in top] == [3, 0]
    assert top[0][0] == pytest.approx(2.1209246397136083, rel=1e-15)


@pytest.mark.parametrize("oracle_cls", [OracleSearcher, NumpyOracle])
def test_zero_length_byte_and_ties(oracle_cls):
    # a document without the field keeps length byte 0 and scores with fl = 1 (W5)
    docs = [{"f": ["a", "q"]}, {"f": ["a", "r"]}, {"f": ["a", "s"]}, {"g": ["a"]}]
    ix = FlatIndex.from_documents(docs, ["f", "g"])
    assert ix.len_bytes[0].tolist() == [2, 2, 2, 0]
    ix.len_bytes[0, 1] = 0          # force a zero byte on a doc that *has* postings
    o = oracle_cls(ix)
    top, total = o.search(Term("f", "a"), limit=2)
    assert total == 3
    idf, avgfl = o.idf("f", "a"), 6 / 4
    want1 = idf * (1 * 2.2) / (1 + 1.2 * (0.25 + 0.75 * 1 / avgfl))
    want0 = idf * (1 * 2.2) / (1 + 1.2 * (0.25 + 0.75 * 2 / avgfl))
    assert top[0] == (pytest.approx(want1, rel=1e-15), 1)
    # docs 0 and 2 tie exactly: lowest docnum is kept at the cut-off (W11)
    assert top[1] == (pytest.approx(want0, rel=1e-15), 0)


@pytest.mark.parametrize("oracle_cls", [OracleSearcher, NumpyOracle])
def test_two_fields_per_field_B(oracle_cls):
    docs = [{"title": ["a"], "body": ["a", "b", "c", "d"]},
            {"title": ["b", "x"], "body": ["a", "a"]},
            {"title": ["a", "a", "a"], "body": ["z"] * 9}]
    ix = FlatIndex.from_documents(docs, ["title", "body"])
    o = oracle_cls(ix, B=0.75, K1=1.2, field_B={"title": 0.3})
    q = Or([Term("title", "a", boost=2.0), Term("body", "a")])
    top, total = o.search(q)
    assert total == 3
    want = {}
    for d, (tt, tb) in enumerate([(1, 1), (0, 2), (3, 0)]):
        s = 0.0
        if tt:
            fl = [1, 2, 3][d]
            s += 2.0 * (o.idf("title", "a") * (tt * 2.2) / (tt + 1.2 * (0.7 + 0.3 * fl / (6 / 3))))
        if tb:
            fl = [4, 2, 9][d]
            s += o.idf("body", "a") * (tb * 2.2) / (tb + 1.2 * (0.25 + 0.75 * fl / (15 / 3)))
        want[d] = s
    assert {d: pytest.approx(s, rel=1e-14) for s, d in top} == want


def test_every_and_clamp():
    docs = [{"f": ["a"]}, {"g": ["a"]}, {"f": ["b"] * 3}]
    ix = FlatIndex.from_documents(docs, ["f", "g"])
    for o in (OracleSearcher(ix), NumpyOracle(ix)):
        top, total = o.search(Every("f"), limit=None)
        assert total == 2 and top == [(1.0, 0), (1.0, 2)]
    assert length_to_byte(200000) == 255


def test_page_view():
    assert page_view(25, 1, 10) == (1, 0, 10, 3)
    assert page_view(25, 3, 10) == (3, 20, 5, 3)
    assert page_view(25, 9, 10) == (3, 20, 5, 3)      # pagenum clamped to the page count
    with pytest.raises(ValueError):
        page_view(25, 0, 10)
