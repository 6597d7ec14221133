"""Whoosh index -> FlatIndex through the reader's public API, driven by a stand-in reader."""
import numpy as np
import pytest

from document_search_engine_b200 import And, FlatIndex, Not, Or, Term
from document_search_engine_b200.flatten import flatten_reader
from document_search_engine_b200.numeric import B2L
from oracle.numpy_oracle import NumpyOracle
from oracle.whoosh_port import OracleSearcher


class _FieldType:
    def __init__(self, scorable):
        self.scorable = scorable

    def from_bytes(self, b):
        return b.decode("utf-8")


class _Matcher:
    def __init__(self, ids, weights, positions=None):
        self.ids, self.w, self.i = ids, weights, 0
        self.positions = positions            # per posting: the word's positions in that document (Whoosh "positions")

    def supports(self, astype):
        return astype == "positions" and self.positions is not None

    def value_as(self, astype):
        assert astype == "positions"
        return self.positions[self.i]

    def is_active(self):
        return self.i < len(self.ids)

    def id(self):
        return int(self.ids[self.i])

    def weight(self):
        return float(self.w[self.i])

    def next(self):
        self.i += 1


class FakeWhooshReader:
    """The part of whoosh.reading.IndexReader that flatten_reader uses, over a FlatIndex (quantised lengths are
    handed back as the decoded lengths, as Whoosh's length column does)."""

    def __init__(self, ix, stored_df=None):
        self.ix = ix
        self.schema = {n: _FieldType(ix.scorable[f]) for f, n in enumerate(ix.field_names)}
        self.stored_df = stored_df

    def indexed_field_names(self):
        return list(self.ix.field_names)

    def lexicon(self, name):
        f = self.ix.field_names.index(name)
        return sorted(t.encode("utf-8") for (ff, t) in self.ix.terms if ff == f)

    def _tid(self, name, btext):
        return self.ix.term_id(name, btext.decode("utf-8"))

    def doc_frequency(self, name, btext):
        tid = self._tid(name, btext)
        return int(self.ix.df[tid]) if self.stored_df is None else int(self.stored_df[tid])

    def postings(self, name, btext):
        tid = self._tid(name, btext)
        d, w = self.ix.postings(tid)
        f = self.ix.field_names.index(name)
        pos = None
        if f in self.ix.positions:
            offs, ids = self.ix.positions[f]
            pos = [np.nonzero(ids[int(offs[x]):int(offs[x + 1])] == tid)[0].tolist() for x in d.tolist()]
        return _Matcher(d, w, pos)

    def doc_field_length(self, docnum, name, default=0):
        b = int(self.ix.len_bytes[self.ix.field_names.index(name), docnum])
        return int(B2L[b]) if b else default

    def field_length(self, name):
        return self.ix.field_length(name)

    def doc_count_all(self):
        return self.ix.n_docs_all

    def is_deleted(self, docnum):
        return self.ix.deleted is not None and bool(self.ix.deleted[docnum])

    def stored_fields(self, docnum):
        return self.ix.stored_fields(docnum)


def _source():
    rng = np.random.default_rng(3)
    docs = []
    for d in range(300):
        toks = ["w%d" % t for t in rng.zipf(1.4, size=int(rng.integers(1, 80))) if t < 150] or ["w1"]
        docs.append({"body": toks, "heading": ["h%d" % (d % 7), "w1"], "book": ["ss", "nopr", "tes1"][d % 3], "session": "s%d" % d})
    return FlatIndex.from_documents(docs, ["body", "heading", "book"], id_fields=["book"], stored=["session", "book"],
                                    deleted=[3, 4, 200])


def test_round_trip_through_the_reader_api():
    src = _source()
    flat = flatten_reader(FakeWhooshReader(src))
    assert flat.field_names == src.field_names and flat.scorable == src.scorable
    assert flat.n_docs_all == src.n_docs_all and flat.doc_count() == src.doc_count()
    assert np.array_equal(flat.field_length_total, src.field_length_total)
    # quantised lengths survive the decode -> re-quantise round trip (W6: the table is its own fixed point)
    assert np.array_equal(flat.len_bytes, src.len_bytes)
    assert np.array_equal(flat.deleted, src.deleted)
    for (f, t), tid in src.terms.items():
        ftid = flat.term_id(src.field_names[f], t)
        assert ftid >= 0 and flat.df[ftid] == src.df[tid]
        d0, w0 = src.postings(tid)
        d1, w1 = flat.postings(ftid)
        assert np.array_equal(d0, d1) and np.array_equal(w0, w1)
    assert flat.stored_fields(7) == src.stored_fields(7)
    # the word order comes back from the per-posting positions: the same documents pass a phrase's positional test
    assert sorted(flat.positions) == sorted(src.positions)
    for words in (["w1", "w2"], ["w1", "w1"], ["w2", "w1", "w3"]):
        assert flat.phrase_docs("body", words).tolist() == src.phrase_docs("body", words).tolist()
    assert src.phrase_docs("body", ["w1", "w1"]).size > 0
    queries = [Term("body", "w1"), And([Term("body", "w2"), Term("heading", "h3")]), Or([Term("body", "w5"), Term("book", "ss")]),
               And([Term("body", "w1"), Not(Term("book", "nopr"))])]
    for cls in (OracleSearcher, NumpyOracle):
        for q in queries:
            assert cls(flat).search(q, limit=20) == cls(src).search(q, limit=20)


def test_stored_document_frequencies_are_kept():
    """The reference deletes and re-adds every document (my_index.py:115-117): until segments merge, the stored df
    and doc_count_all still count the deleted copies (W3), and idf must use them."""
    src = _source()
    stale = src.df.astype(np.int64) * 2
    flat = flatten_reader(FakeWhooshReader(src, stored_df=stale), fields=["body"], stored=False)
    assert flat.field_names == ["body"] and flat.stored is None
    tid = flat.term_id("body", "w1")
    assert flat.df[tid] == stale[src.term_id("body", "w1")] and flat.df[tid] != np.diff(flat.term_offsets)[tid]
    assert NumpyOracle(flat).idf("body", "w1") != NumpyOracle(src).idf("body", "w1")
