"""Parity gate shared by the GPU tests (SURVEY.md §8 d, "Parity gate")."""
import numpy as np

REL_TOL = 1e-5     # north star: GPU fp32 vs Whoosh float64, 1e-5 relative


def assert_query_parity(oracle, q, got_top, got_total, limit, ctx="", abs_tol=None):
    """``got_top``: list of (score, docnum) from the engine; compares with the oracle.

    * total (exact match count) bit-exact;
    * number of hits identical;
    * every returned doc is a true match and its score is within REL_TOL of float64;
    * rank by rank the doc is the oracle's doc, or a doc whose oracle score ties with the
      oracle's score at that rank inside REL_TOL (order identical except among ties).
    """
    d, s = oracle.match_all(q)
    assert got_total == d.size, "%s total %d != oracle %d" % (ctx, got_total, d.size)
    order = np.lexsort((d, -s))
    want_n = d.size if limit is None else min(limit, d.size)
    assert len(got_top) == want_n, "%s hits %d != %d" % (ctx, len(got_top), want_n)
    if want_n == 0:
        return
    score_of = dict(zip(d.tolist(), s.tolist()))

    def close(a, b):
        if abs_tol is None:
            return abs(a - b) <= REL_TOL * abs(b)
        # final() values (tests/test_date_final.py).  Dated documents (> 1): differences live in the last few
        # ulps, abs_tol.  Undated ones are 1 - 1/s: a relative error r in s moves them by r / s = r * (1 - v).
        return abs(a - b) <= (abs_tol if b > 1.0 else REL_TOL * abs(1.0 - b))

    seen = set()
    prev = None
    for i, (gs, gd) in enumerate(got_top):
        assert gd in score_of, "%s rank %d: doc %d is not a match" % (ctx, i, gd)
        assert gd not in seen, "%s rank %d: doc %d returned twice" % (ctx, i, gd)
        seen.add(gd)
        ws = score_of[gd]
        assert close(gs, ws), "%s rank %d doc %d: score %r vs %r" % (ctx, i, gd, gs, ws)
        od, os_ = int(d[order[i]]), float(s[order[i]])
        if gd != od:
            assert close(ws, os_), \
                "%s rank %d: doc %d (%.9g) where oracle has doc %d (%.9g)" % (ctx, i, gd, ws, od, os_)
        if prev is not None:       # engine's own order: score desc, docnum asc
            assert (gs < prev[0]) or (gs == prev[0] and gd > prev[1]), "%s rank %d out of order" % (ctx, i)
        prev = (gs, gd)


def assert_batch_parity(oracle, queries, results, limit, sample=None, abs_tol=None):
    idx = range(len(queries)) if sample is None else sample
    for i in idx:
        r = results[i]
        assert_query_parity(oracle, queries[i], r.top_n, len(r), limit, ctx="query %d %s:" % (i, queries[i]), abs_tol=abs_tol)
