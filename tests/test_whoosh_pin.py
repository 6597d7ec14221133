"""Pin-on-arrival: the oracles (and through them every parity test) against a real Whoosh 2.7.4 searcher.

Whoosh (reference ``requirements.txt:6``) is not in this image, so the module is skipped here; it runs wherever
``import whoosh`` works or a driver-provided install sits in ``baseline/_ref`` (SURVEY.md
section 8 d).  The synthetic corpus is written as ``t0000123`` tokens and indexed with a whitespace tokenizer only,
so term ids map 1:1; ``searcher.search(q, limit=k)`` with ``weighting=BM25F`` is the reference's own call
(``my_flask.py:183-184``, ``:208``, ``:211``, ``:304``).  What must hold: identical matched-document sets and totals,
identical order, scores within 1e-12 relative (float64 both sides), for Term / And / Or / And-of-Or / Not, a
non-scorable ID field (W15), deleted documents (W9), per-field B and leaf boosts, and for the committed golden
fixtures.
"""
import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(_REF) and _REF not in sys.path:
    sys.path.append(_REF)
whoosh = pytest.importorskip("whoosh", reason="Whoosh is not installed here: parity stays unpinned (oracle/whoosh_port.py header)")

from document_search_engine_b200 import And, FlatIndex, Not, Or, Term       # noqa: E402
from document_search_engine_b200.corpus import make_corpus, make_queries      # noqa: E402
from oracle.numpy_oracle import NumpyOracle                                   # noqa: E402
from oracle.whoosh_port import OracleSearcher                                 # noqa: E402


def whoosh_index(tmp_path, docs, id_fields=(), deleted=(), phrase=False):
    """``docs``: list of {field: token list | id string}; one commit = one segment, docnum = position."""
    from whoosh import fields, index
    from whoosh.analysis import SpaceSeparatedTokenizer
    names = sorted({k for d in docs for k in d})
    schema = fields.Schema(**{n: (fields.ID() if n in id_fields else fields.TEXT(analyzer=SpaceSeparatedTokenizer(), phrase=phrase))
                              for n in names})
    ix = index.create_in(str(tmp_path), schema)
    w = ix.writer()
    for d in docs:
        w.add_document(**{k: (v if isinstance(v, str) else " ".join(v)) for k, v in d.items()})
    w.commit()
    if deleted:
        w = ix.writer()
        for dn in deleted:
            w.delete_document(dn)
        w.commit(merge=False)           # keep the segment: df and doc_count_all still count the deleted documents (W3, W9)
    return ix


def to_whoosh(q):
    from whoosh import query as wq
    name = type(q).__name__
    if name == "Term":
        return wq.Term(q.fieldname, q.text if isinstance(q.text, str) else "t%07d" % q.text, boost=q.boost)
    if name == "Not":
        return wq.Not(to_whoosh(q.query))
    if name == "Phrase":
        return wq.Phrase(q.fieldname, list(q.words), slop=q.slop, boost=q.boost)
    cls = {"And": wq.And, "Or": wq.Or}[name]
    return cls([to_whoosh(s) for s in q.subqueries], boost=q.boost)


def whoosh_search(ix, q, limit, **bm25f_kwargs):
    from whoosh import scoring
    with ix.searcher(weighting=scoring.BM25F(**bm25f_kwargs)) as s:
        r = s.search(to_whoosh(q), limit=limit)
        return [(hit.score, hit.docnum) for hit in r], len(r)


def assert_same(got, want, ctx):
    (gt, gn), (wt, wn) = got, want
    assert gn == wn, "%s: total %d vs Whoosh %d" % (ctx, gn, wn)
    assert [d for _, d in gt] == [d for _, d in wt], "%s: order differs" % ctx
    for (gs, d), (ws, _) in zip(gt, wt):
        assert gs == pytest.approx(ws, rel=1e-12), "%s doc %d: %r vs Whoosh %r" % (ctx, d, gs, ws)


def token_docs(ix):
    """The documents of a generated FlatIndex as token lists (term rank r -> 't%07d')."""
    docs = [dict() for _ in range(ix.n_docs_all)]
    for tid in range(ix.n_terms):
        f = ix.field_names[int(ix.term_field[tid])]
        r = tid - int(ix.term_field[tid]) * ix.vocab_size
        d, tf = ix.postings(tid)
        for dn, n in zip(d.tolist(), tf.tolist()):
            docs[dn].setdefault(f, []).extend(["t%07d" % r] * int(n))
    return docs


def test_synthetic_corpus_matches_whoosh(tmp_path):
    ix = make_corpus(2000, 5000, 20260001, device="cpu")
    wix = whoosh_index(tmp_path, token_docs(ix))
    qs = make_queries(200, 5000, 20261001, 1, 4, "mixed", skip_top=10).queries
    qs += make_queries(60, 5000, 31, 4, 4, "and", variants=True, skip_top=0).queries
    for o in (OracleSearcher(ix), NumpyOracle(ix)):
        for i, q in enumerate(qs):
            for k in (3, 10, 150, None):
                assert_same(o.search(q, limit=k), whoosh_search(wix, q, k), "%s query %d %s limit %s" % (type(o).__name__, i, q, k))


def test_small_cases_match_whoosh(tmp_path):
    docs = [{"body": "seth speaks of joy".split(), "book": "ss"}, {"body": "joy and vitality joy".split(), "book": "nopr"},
            {"body": "the nature of joy".split(), "book": "nopr"}, {"body": ["dreams"], "book": "deavf1"},
            {"body": ["joy"] * 40 + ["x%d" % i for i in range(300)], "book": "ss"}]
    fix = FlatIndex.from_documents(docs, ["body", "book"], id_fields=["book"], deleted=[2])
    wix = whoosh_index(tmp_path, docs, id_fields=["book"], deleted=[2])
    queries = [Term("body", "joy"), Term("book", "nopr"), Term("book", "nopr", boost=2.5), Term("body", "nope"),
               And([Term("body", "joy"), Term("book", "ss")]), And([Term("body", "joy"), Not(Term("book", "ss"))]),
               Or([Term("body", "joy", boost=0.5), Term("body", "dreams"), Term("book", "deavf1")]),
               And([Or([Term("body", "joy"), Term("body", "dreams")]), Or([Term("book", "ss"), Term("book", "deavf1")])])]
    for kw in ({}, {"B": 0.3, "K1": 2.0}, {"body_B": 0.1}):
        okw = {"B": kw.get("B", 0.75), "K1": kw.get("K1", 1.2), "field_B": {"body": kw["body_B"]} if "body_B" in kw else None}
        for o in (OracleSearcher(fix, **okw), NumpyOracle(fix, **okw)):
            for q in queries:
                assert_same(o.search(q, limit=10), whoosh_search(wix, q, 10, **kw), "%s %s %s" % (type(o).__name__, kw, q))


def test_phrases_match_whoosh(tmp_path):
    """query.Phrase (the reference UI's quoted phrases, search-form.html:20-40): which documents pass the positional
    test, and that they score the sum of their words' scores."""
    from document_search_engine_b200 import Phrase
    from tests.test_phrases import corpus, phrases
    fix = corpus(300, seed=5)
    offs, ids = fix.positions[0]
    text_of = {tid: t for (f, t), tid in fix.terms.items()}
    docs = []
    for d in range(fix.n_docs_all):
        doc = {}
        for f, name in enumerate(fix.field_names):
            o, i = fix.positions[f]
            toks = [text_of[int(t)] for t in i[int(o[d]):int(o[d + 1])]]
            if toks:
                doc[name] = toks
        docs.append(doc)
    deleted = np.nonzero(fix.deleted)[0].tolist()
    wix = whoosh_index(tmp_path, docs, deleted=deleted, phrase=True)
    qs = [p for p in phrases() if p.fieldname in fix.field_names] + [
        And([Term("body", "w0"), Phrase("body", ["w1", "w2"])]), And([Term("body", "w0"), Not(Phrase("body", ["w1", "w2"]))])]
    for o in (OracleSearcher(fix), NumpyOracle(fix)):
        for q in qs:
            assert_same(o.search(q, limit=20), whoosh_search(wix, q, 20), "%s %s" % (type(o).__name__, q))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.json"))))
def test_golden_fixtures_match_whoosh(tmp_path, path):
    """The committed fixtures were written by the doc-at-a-time port (tests/golden/make_golden.py); with Whoosh at
    hand they are re-derived from it."""
    from tests.golden.make_golden import build_case
    with open(path) as f:
        fx = json.load(f)
    c = fx["case"]
    ix, queries = build_case(c)
    if ix.terms is not None or any(type(q).__name__ == "Every" for q in queries):
        pytest.skip("string vocabulary / Every(): covered by test_small_cases_match_whoosh and the CLI harness")
    deleted = [] if ix.deleted is None else np.nonzero(ix.deleted)[0].tolist()
    wix = whoosh_index(tmp_path, token_docs(ix), deleted=deleted)
    kw = dict(B=c["B"], K1=c["K1"], **{f + "_B": b for f, b in c["field_B"].items()})
    for q, want in zip(queries, fx["results"]):
        got_top, got_total = whoosh_search(wix, q, c["k"], **kw)
        assert got_total == want["total"] and [d for _, d in got_top] == [d for _, d in want["top"]]
        assert [s for s, _ in got_top] == pytest.approx([s for s, _ in want["top"]], rel=1e-12)


def test_flattener_on_a_real_index(tmp_path):
    """flatten_index over Whoosh's own reader gives the arrays FlatIndex.from_documents builds from the same tokens."""
    from document_search_engine_b200.flatten import flatten_index
    docs = [{"body": ("w%d w%d w1" % (d % 13, d % 7)).split() * (1 + d % 3), "book": ["ss", "nopr", "tes1"][d % 3]} for d in range(120)]
    wix = whoosh_index(tmp_path, docs, id_fields=["book"], deleted=[5, 6])
    flat = flatten_index(wix, fields=["body", "book"])
    ref = FlatIndex.from_documents(docs, ["body", "book"], id_fields=["book"], deleted=[5, 6])
    assert flat.scorable == ref.scorable and np.array_equal(flat.len_bytes, ref.len_bytes)
    assert np.array_equal(flat.field_length_total, ref.field_length_total) and np.array_equal(flat.deleted, ref.deleted)
    for (f, t), tid in ref.terms.items():
        ftid = flat.term_id(ref.field_names[f], t)
        assert ftid >= 0 and flat.df[ftid] == ref.df[tid]
        assert np.array_equal(flat.postings(ftid)[0], ref.postings(tid)[0]) and np.array_equal(flat.postings(ftid)[1], ref.postings(tid)[1])
    for q in (Term("body", "w1"), And([Term("body", "w2"), Not(Term("book", "ss"))]), Or([Term("body", "w3"), Term("book", "tes1")])):
        assert_same(OracleSearcher(flat).search(q, limit=10), whoosh_search(wix, q, 10), "flattened %s" % q)
