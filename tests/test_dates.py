"""Date ranges (reference UI: ``date:[oct 1970 to dec 8 1970]``, ``hospital date:"feb 1964"``, search-form.html:26, :39):
the year / month / day tokens of the flat index, the range cover, the parser, and - on the CPU - that the expansion
into posting lists selects and scores exactly what the oracles compute from the stored dates."""
import random
from datetime import date, datetime, timedelta

import numpy as np
import pytest

from document_search_engine_b200 import And, DateRange, FlatIndex, Not, Or, QueryParser, Term, dates
from document_search_engine_b200.query import NullQuery, expand_multiterms
from oracle.numpy_oracle import NumpyOracle
from oracle.whoosh_port import OracleSearcher


def days_of(tokens):
    out = set()
    for t in tokens:
        if t[0] == "Y":
            d = date(int(t[1:]), 1, 1)
            while d.year == int(t[1:]):
                out.add(d)
                d += timedelta(days=1)
        elif t[0] == "M":
            y, m = int(t[1:5]), int(t[6:8])
            d = date(y, m, 1)
            while d.month == m:
                out.add(d)
                d += timedelta(days=1)
        else:
            out.add(date(int(t[1:5]), int(t[6:8]), int(t[9:11])))
    return out


def test_range_cover_is_exact_and_short():
    rng = random.Random(7)
    for _ in range(400):
        a = date(1960, 1, 1) + timedelta(days=rng.randrange(0, 9000))
        b = a + timedelta(days=rng.randrange(0, 2500))
        cover = dates.range_cover(a, b)
        assert days_of(cover) == {a + timedelta(days=i) for i in range((b - a).days + 1)}
        assert len(cover) == len(set(cover)) <= 2 * (30 + 11) + (b.year - a.year + 1)
    assert dates.range_cover(date(1970, 1, 1), date(1971, 12, 31)) == ["Y1970", "Y1971"]
    assert dates.range_cover(date(1970, 10, 1), date(1970, 12, 8)) == ["M1970-10", "M1970-11"] + ["D1970-12-%02d" % d for d in range(1, 9)]
    assert dates.range_cover(date(1970, 3, 2), date(1970, 3, 1)) == []


def test_day_bounds_of_a_range():
    assert dates.first_day(datetime(1970, 10, 1)) == date(1970, 10, 1)
    assert dates.first_day(datetime(1970, 10, 1, 0, 0, 1)) == date(1970, 10, 2)      # midnight of Oct 1 is before the start
    assert dates.last_day(datetime(1970, 12, 8, 23, 59, 59, 999999)) == date(1970, 12, 8)
    assert dates.last_day(datetime(1970, 12, 8)) == date(1970, 12, 8)
    with pytest.raises(ValueError):
        dates.tier_tokens(datetime(1970, 1, 1, 12))


def test_date_expressions():
    assert dates.parse_span("feb 1964") == (datetime(1964, 2, 1), datetime(1964, 2, 29, 23, 59, 59, 999999))
    assert dates.parse_span("1964") == (datetime(1964, 1, 1), datetime(1964, 12, 31, 23, 59, 59, 999999))
    assert dates.parse_span("Dec 8, 1970")[0] == dates.parse_span("8 december 1970")[0] == datetime(1970, 12, 8)
    assert dates.parse_span("1970-12-08") == dates.parse_span("19701208") == dates.parse_span("dec 8 1970")
    assert dates.parse_range("oct 1970 to dec 8 1970") == (datetime(1970, 10, 1), datetime(1970, 12, 8, 23, 59, 59, 999999))
    assert dates.parse_range("to 1965") == (None, datetime(1965, 12, 31, 23, 59, 59, 999999))
    assert dates.parse_range("1965 to") == (datetime(1965, 1, 1), None)
    for bad in ("soon", "feb", "31 feb 1970", "8 1970"):
        with pytest.raises(dates.DateParseError):
            dates.parse_span(bad)


def test_parser_builds_date_ranges():
    qp = QueryParser("body")
    q = qp.parse('hospital date:"feb 1964"')
    assert q == And([Term("body", "hospital"), DateRange("date", datetime(1964, 2, 1), datetime(1964, 2, 29, 23, 59, 59, 999999))])
    q = qp.parse("date:[oct 1970 to dec 8 1970]")
    assert isinstance(q, DateRange) and q.start == datetime(1970, 10, 1) and q.end.date() == date(1970, 12, 8)
    q = qp.parse("seth date:1970 NOT jane")
    assert [type(s).__name__ for s in q.subqueries] == ["Term", "DateRange", "Not"]
    assert qp.parse('exact:"feb 1964"') != q                      # only date fields get the date grammar
    with pytest.raises(dates.DateParseError):                      # the reference catches this and redirects (my_flask.py:193-196)
        qp.parse("date:[whenever to 1970]")


def corpus(n=600, seed=3):
    rng = random.Random(seed)
    words = ["w%d" % i for i in range(40)]
    docs = []
    for d in range(n):
        doc = {"body": [rng.choice(words) for _ in range(rng.randrange(3, 30))]}
        if rng.random() < 0.9:
            day = date(1963, 6, 1) + timedelta(days=rng.randrange(0, 3000))
            doc["date"] = datetime(day.year, day.month, day.day) if rng.random() < 0.5 else day
        docs.append(doc)
    return FlatIndex.from_documents(docs, ["body"], stored=["date"], date_fields=["date"], deleted=[d for d in (5, 17, 300) if d < n])


def queries():
    return [DateRange("date", datetime(1964, 2, 1), datetime(1964, 2, 29, 23, 59, 59, 999999)),
            DateRange("date", date(1965, 3, 7), date(1969, 11, 20)),
            DateRange("date", None, datetime(1964, 1, 1)),
            DateRange("date", datetime(1970, 1, 1, 0, 0, 1), None),
            DateRange("date", datetime(1990, 1, 1), datetime(1991, 1, 1)),                      # nothing there
            DateRange("date", datetime(1966, 5, 5), datetime(1966, 5, 5), boost=2.5),          # one day
            And([Term("body", "w3"), DateRange("date", datetime(1966, 1, 1), datetime(1968, 6, 30))]),
            And([Or([Term("body", "w1"), Term("body", "w2")]), DateRange("date", date(1964, 1, 15), date(1964, 9, 2), boost=3.0)]),
            Or([Term("body", "w7"), DateRange("date", date(1967, 1, 1), date(1967, 12, 31))]),
            And([Term("body", "w4"), Not(DateRange("date", date(1964, 1, 1), date(1969, 12, 31)))])]


def test_expansion_selects_and_scores_what_the_oracles_compute_from_stored_dates():
    ix = corpus()
    assert not ix.is_scorable("date") and ix.is_scorable("body")
    no, wo = NumpyOracle(ix), OracleSearcher(ix)
    for q in queries():
        d, s = no.match_all(q)                                   # brute force over the stored dates
        top, total = wo.search(q, limit=None)
        assert total == d.size and sorted(doc for _, doc in top) == d.tolist()
        byd = dict(zip(d.tolist(), s.tolist()))
        assert all(abs(sc - byd[doc]) <= 1e-12 * abs(sc) for sc, doc in top)
        e = expand_multiterms(q, ix.lexicon)                    # what the engine lowers: posting lists of date tokens
        if e is NullQuery or (hasattr(e, "subqueries") and any(x is NullQuery for x in e.subqueries)):
            assert d.size == 0 or type(q) is not DateRange
            continue
        d2, s2 = no.match_all(e.normalize() if hasattr(e, "normalize") else e)
        assert d2.tolist() == d.tolist(), q
        assert np.allclose(s2, s, rtol=1e-12, atol=0), q


def test_date_tokens_survive_save_load_and_sharding(tmp_path):
    ix = corpus(200, seed=9)
    p = str(tmp_path / "ix.npz")
    ix.save(p)
    back = FlatIndex.load(p)
    assert back.lexicon("date") == ix.lexicon("date") and not back.is_scorable("date")
    q = DateRange("date", date(1965, 1, 1), date(1966, 12, 31))
    whole = NumpyOracle(ix).match_all(expand_multiterms(q, ix.lexicon))[0]
    parts = NumpyOracle(ix, shards=[ix.shard(g, 3) for g in range(3)]).match_all(expand_multiterms(q, ix.lexicon))[0]
    assert sorted(parts.tolist()) == whole.tolist()
