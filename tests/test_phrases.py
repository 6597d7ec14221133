"""Phrase queries (reference UI: quoted phrases, search-form.html:20-40; Whoosh query.Phrase): word order in the flat
index, the positional test on the host, the rewrite into And(words, per-batch document list), the parser, and the two
oracles' restatement of Whoosh's semantics (the score is the words' alone; the positions only decide the match)."""
import random

import numpy as np
import pytest

from document_search_engine_b200 import And, FlatIndex, Not, Or, Phrase, QueryParser, Term
from document_search_engine_b200.query import FILTER_FIELD, NullQuery, expand_phrases, has_phrase, lower
from oracle import numpy_oracle
from oracle.numpy_oracle import NumpyOracle
from oracle.whoosh_port import OracleSearcher


def corpus(n=400, seed=5, vocab=12):
    rng = random.Random(seed)
    words = ["w%d" % i for i in range(vocab)]
    docs = [{"body": [rng.choice(words) for _ in range(rng.randrange(0, 40))],
             "title": [rng.choice(words) for _ in range(rng.randrange(0, 5))]} for _ in range(n)]
    return FlatIndex.from_documents(docs, ["body", "title"], deleted=[d for d in (3, 44) if d < n])


def phrases():
    return [Phrase("body", ["w1", "w2"]), Phrase("body", ["w3", "w3"]), Phrase("body", ["w1", "w2", "w3"]),
            Phrase("body", ["w4", "w5"], slop=3), Phrase("body", ["w6", "w7", "w8"], slop=2, boost=2.0),
            Phrase("title", ["w1", "w2"]), Phrase("body", ["w1", "nope"]), Phrase("body", ["w9"]), Phrase("nofield", ["a", "b"])]


def test_positional_test_against_a_plain_scan():
    ix = corpus()
    for p in phrases():
        if len(p.words) < 2:
            continue
        want = numpy_oracle.phrase_docs(ix, p.fieldname, p.words, p.slop)
        got = ix.phrase_docs(p.fieldname, p.words, p.slop)
        assert got.tolist() == want.tolist(), p
    assert ix.phrase_docs("body", ["w1", "w2"]).size > 0
    # slop widens, never narrows
    assert set(ix.phrase_docs("body", ["w4", "w5"]).tolist()) <= set(ix.phrase_docs("body", ["w4", "w5"], slop=3).tolist())


def test_word_order_survives_sharding_and_save_load(tmp_path):
    ix = corpus(150, seed=8)
    whole = ix.phrase_docs("body", ["w1", "w2"]).tolist()
    parts = []
    for g in range(3):
        sh = ix.shard(g, 3)
        parts += [int(d) + sh.doc_base for d in sh.phrase_docs("body", ["w1", "w2"])]
    assert parts == whole
    p = str(tmp_path / "ix.npz")
    ix.save(p)
    assert FlatIndex.load(p).phrase_docs("body", ["w1", "w2"]).tolist() == whole


def test_oracles_agree_on_phrases():
    ix = corpus()
    no, wo = NumpyOracle(ix), OracleSearcher(ix)
    qs = phrases() + [And([Term("body", "w0"), Phrase("body", ["w1", "w2"])]),
                      And([Phrase("body", ["w1", "w2"]), Phrase("body", ["w2", "w1"], boost=3.0)]),
                      And([Term("body", "w0"), Not(Phrase("body", ["w1", "w2"]))])]
    for q in qs:
        d, s = no.match_all(q)
        top, total = wo.search(q, limit=None)
        assert total == d.size, q
        assert sorted(doc for _, doc in top) == d.tolist(), q
        byd = dict(zip(d.tolist(), s.tolist()))
        assert all(abs(sc - byd[doc]) <= 1e-12 * abs(sc) for sc, doc in top), q
    # a phrase scores exactly what the AND of its words scores on the documents it lets through
    d, s = no.match_all(Phrase("body", ["w1", "w2"]))
    da, sa = no.match_all(And([Term("body", "w1"), Term("body", "w2")]))
    assert set(d.tolist()) < set(da.tolist())
    assert np.array_equal(s, sa[np.isin(da, d)])


def test_rewrite_and_parser():
    reg = []

    def register(p):
        if p not in reg:
            reg.append(p)
        return reg.index(p)
    q = And([Term("body", "a"), Phrase("body", ["b", "c"], boost=2.0), Not(Phrase("body", ["d", "e"]))])
    assert has_phrase(q) and not has_phrase(And([Term("body", "a")]))
    leaves, g, kind = lower(expand_phrases(q, register))
    assert kind == "groups" and g == 4
    assert [(lf.fieldname, lf.text, lf.boost, lf.group) for lf in leaves] == [
        ("body", "a", 1.0, 0), ("body", "b", 2.0, 1), ("body", "c", 2.0, 2), (FILTER_FIELD, 0, 2.0, 3), (FILTER_FIELD, 1, 1.0, 255)]
    assert reg == [Phrase("body", ["b", "c"], boost=2.0), Phrase("body", ["d", "e"])]
    assert expand_phrases(Phrase("body", ["x"]), register) == Term("body", "x") and expand_phrases(Phrase("body", []), register) is NullQuery
    qp = QueryParser("body")
    assert qp.parse('seth "Dead Sea" jane') == And([Term("body", "seth"), Phrase("body", ["dead", "sea"]), Term("body", "jane")])
    assert qp.parse('title:"way toward health"~3') == Phrase("title", ["way", "toward", "health"], slop=3)
    assert qp.parse('"single"') == Term("body", "single")
    assert qp.parse('a NOT "b c"') == And([Term("body", "a"), Not(Phrase("body", ["b", "c"]))])
