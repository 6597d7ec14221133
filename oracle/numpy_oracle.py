"""CPU ORACLE (test infrastructure, not product code) — term-at-a-time, vectorised.

PARITY UNPINNED (see ``oracle/whoosh_port.py`` for why): same W1-W14 restatement
of Whoosh 2.7.4 semantics, evaluated one posting *list* at a time with numpy in
float64 so that million-document ground truth is affordable.  ``tests/`` checks
that this and the doc-at-a-time port agree on random indexes; both are pinned
to KAT-1 of SURVEY.md §8 c.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this module.

Reference call sites this follows: ``my_flask.py:183-184`` (BM25F defaults, W2),
``:208``/``:211``/``:304`` (limits), ``my_index.py:172-177`` (TEXT fields).
"""
from __future__ import annotations

from math import log

import numpy as np

B2L = np.array([int(round((pow(1.033, i) - 1) * 27)) for i in range(256)], dtype=np.float64)
B2L_SCORING = B2L.copy()
B2L_SCORING[0] = 1.0            # W5: byte 0 → default length 1


def _expand(q, ix):
    """Prefix / Wildcard -> the words of the field's lexicon that fit, in lexicon order (Whoosh MultiTerm.matcher)."""
    import fnmatch
    f = ix.field_names.index(q.fieldname) if q.fieldname in ix.field_names else -1
    fits = (lambda w: w.startswith(q.text)) if type(q).__name__ == "Prefix" else (lambda w: fnmatch.fnmatchcase(w, q.text))
    return sorted(t for (ff, t) in (ix.terms or {}) if ff == f and isinstance(t, str) and fits(t))


def date_range_docs(sub, fieldname, start, end):
    """Local docnums of ``sub`` whose stored ``fieldname`` date lies in ``[start, end]`` (``None`` = open), by brute
    force over the stored fields - deliberately not through the index's date tokens, which are what is under test.
    Whoosh: query.DateRange -> NumericRange over the DATETIME field (reference search-form.html:26, :39)."""
    from datetime import datetime
    out = []
    for d, sf in enumerate(sub.stored or []):
        v = sf.get(fieldname)
        if v is None:
            continue
        when = v if isinstance(v, datetime) else datetime(v.year, v.month, v.day)
        if (start is None or _as_dt(start) <= when) and (end is None or when <= _as_dt(end)):
            out.append(d)
    return np.asarray(out, dtype=np.int64)


def _as_dt(v):
    from datetime import datetime
    return v if isinstance(v, datetime) else datetime(v.year, v.month, v.day)


def phrase_docs(sub, fieldname, words, slop):
    """Local docnums of ``sub`` that contain ``words`` in order, each 1 .. ``slop`` positions after the one before
    (Whoosh query.Phrase -> spans.SpanNear chain, ordered; slop 1 = adjacent): a plain scan of every document's token
    sequence (``FlatIndex.positions``), independent of the engine's candidate-driven ``FlatIndex.phrase_docs``."""
    if fieldname not in sub.field_names:
        return np.zeros(0, np.int64)
    f = sub.field_names.index(fieldname)
    tids = [sub.term_id(fieldname, w) for w in words]
    if min(tids) < 0:
        return np.zeros(0, np.int64)
    offs, ids = sub.positions[f]
    out = []
    for d in range(offs.size - 1):
        seq = ids[int(offs[d]):int(offs[d + 1])].tolist()
        ends = [p for p, t in enumerate(seq) if t == tids[0]]
        for t in tids[1:]:
            ends = [p for p, x in enumerate(seq) if x == t and any(1 <= p - e <= slop for e in ends)]
            if not ends:
                break
        if ends:
            out.append(d)
    return np.asarray(out, dtype=np.int64)


def lower_query(q, ix=None):
    """``(groups, negatives, kind)``: ``groups`` is a list of OR-groups, each a list of
    ``(fieldname, text, boost)``; all groups must match (W10) and no leaf of ``negatives``
    (``(fieldname, text)`` from Not children, Whoosh's AndNotMatcher) may."""
    name = type(q).__name__
    if name == "_Null":
        return [], [], "null"
    if name == "Every":
        return [[(q.fieldname, None, q.boost)]], [], "every"
    if name == "Term":
        return [[(q.fieldname, q.text, q.boost)]], [], "groups"
    if name == "Phrase":
        # the words' scores add up (IntersectionMatcher under the span filter); the positional test is one more
        # "group" that scores nothing
        if not q.words:
            return [], [], "null"
        if len(q.words) == 1:
            return [[(q.fieldname, q.words[0], q.boost)]], [], "groups"
        return ([[(q.fieldname, w, q.boost)] for w in q.words] + [[(q.fieldname, ("phrase", tuple(q.words), q.slop), 0.0)]]), [], "groups"
    if name == "DateRange":
        # constant score: every document of the range scores the boost (ConstantScoreQuery)
        return [[(q.fieldname, ("daterange", q.start, q.end), q.boost)]], [], "groups"
    if name in ("Prefix", "Wildcard"):
        words = _expand(q, ix)
        if not words:
            return [], [], "null"
        boost = 1.0 if len(words) == 1 else q.boost               # a single word: its plain term matcher, no boost
        return [[(q.fieldname, w, boost) for w in words]], [], "groups"
    if name in ("Or", "And"):
        groups, flat, neg = [], [], []
        for s in q.subqueries:
            sn = type(s).__name__
            if sn in ("Prefix", "Wildcard"):
                words = _expand(s, ix)
                b = (1.0 if len(words) == 1 else s.boost) * q.boost
                if name == "And":
                    groups.append([(s.fieldname, w, b) for w in words])      # no word: an empty group, nothing matches
                else:
                    flat.extend((s.fieldname, w, b) for w in words)
                continue
            if sn == "Phrase" and name == "And" and len(s.words) > 1:
                groups.extend([(s.fieldname, w, s.boost * q.boost)] for w in s.words)
                groups.append([(s.fieldname, ("phrase", tuple(s.words), s.slop), 0.0)])
                continue
            if sn == "Phrase" and len(s.words) == 1:
                leaf = (s.fieldname, s.words[0], s.boost * q.boost)
                (groups if name == "And" else flat).append([leaf] if name == "And" else leaf)
                continue
            if sn == "DateRange":
                leaf = (s.fieldname, ("daterange", s.start, s.end), s.boost * q.boost)
                (groups if name == "And" else flat).append([leaf] if name == "And" else leaf)
                continue
            if sn == "Term":
                (groups if name == "And" else flat).append(
                    [(s.fieldname, s.text, s.boost * q.boost)] if name == "And" else (s.fieldname, s.text, s.boost * q.boost))
            elif sn == "Or" and name == "And":
                groups.append([(t.fieldname, t.text, t.boost * s.boost * q.boost) for t in s.subqueries])
            elif sn == "Not":
                inner = s.query.subqueries if type(s.query).__name__ == "Or" else [s.query]
                for t in inner:
                    if type(t).__name__ in ("Prefix", "Wildcard"):
                        neg.extend((t.fieldname, w) for w in _expand(t, ix))
                    elif type(t).__name__ == "DateRange":
                        neg.append((t.fieldname, ("daterange", t.start, t.end)))
                    elif type(t).__name__ == "Phrase":
                        neg.append((t.fieldname, ("phrase", tuple(t.words), t.slop) if len(t.words) > 1 else t.words[0]))
                    else:
                        neg.append((t.fieldname, t.text))
            else:
                raise NotImplementedError(s)
        if flat:
            groups = [flat]
        if not groups:
            return [], [], "null"             # Whoosh: a compound query without positive subqueries matches nothing
        return groups, neg, "groups"
    raise NotImplementedError(name)


class NumpyOracle:
    def __init__(self, ix, B=0.75, K1=1.2, field_B=None, shards=None, final_add=None):
        self.ix = ix
        #: W14 with the reference's DateBM25F.final (my_whoosh.py:129-146), vectorised: per global docnum
        #: ``date seconds + 1.0`` or NaN (no date); every match's score s becomes 1 - 1/s, and
        #: ((1 - 1/s) + final_add) / 10**9 for a dated document, before ranking
        self.final_add = None if final_add is None else np.asarray(final_add, dtype=np.float64)
        self.B, self.K1 = B, K1
        self.field_B = dict(field_B or {})
        self.shards = list(shards) if shards is not None else [ix]

    def idf(self, fieldname, text):
        tid = self.ix.term_id(fieldname, text)
        df = 0 if tid < 0 else int(self.ix.df[tid])
        return log(self.ix.doc_count_all() / (df + 1)) + 1                      # W3

    def avgfl(self, fieldname):
        f = self.ix.field_names.index(fieldname)
        return (int(self.ix.field_length_total[f]) / (self.ix.doc_count_all() or 1)) or 1   # W4

    def leaf_scores(self, sub, fieldname, text, boost):
        """(local docids, float64 scores) of one leaf in one shard; deleted docs removed (W9)."""
        if isinstance(text, tuple) and text and text[0] == "phrase":
            d = phrase_docs(sub, fieldname, text[1], text[2])
            if sub.deleted is not None:
                d = d[sub.deleted[d] == 0]
            return d, np.zeros(d.size, np.float64)
        if isinstance(text, tuple) and text and text[0] == "daterange":
            d = date_range_docs(sub, fieldname, text[1], text[2])
            if sub.deleted is not None:
                d = d[sub.deleted[d] == 0]
            return d, np.full(d.size, float(boost))
        tid = sub.term_id(fieldname, text)
        if tid < 0:
            return np.zeros(0, np.int64), np.zeros(0, np.float64)
        a, b = int(sub.term_offsets[tid]), int(sub.term_offsets[tid + 1])
        d = sub.docids[a:b].astype(np.int64)
        tf = sub.tfs[a:b].astype(np.float64)
        f = sub.field_names.index(fieldname)
        flags = getattr(sub, "scorable", None)
        if flags is not None and not flags[f]:
            s = tf.copy()                                                         # W15: WeightScorer, the posting weight
        else:
            fl = B2L_SCORING[sub.len_bytes[f][d]]
            B = self.field_B.get(fieldname, self.B)
            K1 = self.K1
            s = self.idf(fieldname, text) * ((tf * (K1 + 1)) / (tf + K1 * ((1 - B) + B * fl / self.avgfl(fieldname))))  # W1
        if boost != 1.0:
            s = s * boost
        if sub.deleted is not None:
            keep = sub.deleted[d] == 0
            d, s = d[keep], s[keep]
        return d, s

    def match_all(self, q):
        """All matches as (global docids ascending, float64 scores)."""
        groups, negatives, kind = lower_query(q, self.ix)
        if kind == "null":
            return np.zeros(0, np.int64), np.zeros(0, np.float64)
        ds, ss = [], []
        for sub in self.shards:
            n = sub.len_bytes.shape[1]
            if kind == "every":
                fname, _, boost = groups[0][0]
                if fname not in sub.field_names:
                    continue
                m = sub.len_bytes[sub.field_names.index(fname)] != 0
                if sub.deleted is not None:
                    m &= sub.deleted == 0
                d = np.nonzero(m)[0]
                ds.append(d + sub.doc_base)
                ss.append(np.full(d.size, float(boost)))
                continue
            acc = np.zeros(n, np.float64)
            cnt = np.zeros(n, np.int32)
            for g in groups:
                hit = np.zeros(n, bool)
                for fname, text, boost in g:
                    d, s = self.leaf_scores(sub, fname, text, boost)
                    acc[d] += s                                                  # W10: sum of leaf scores
                    hit[d] = True
                cnt += hit
            for fname, text in negatives:
                d, _ = self.leaf_scores(sub, fname, text, 1.0)
                cnt[d] = -1                                                      # AndNot: excluded
            d = np.nonzero(cnt == len(groups))[0]
            ds.append(d + sub.doc_base)
            ss.append(acc[d])
        if not ds:
            return np.zeros(0, np.int64), np.zeros(0, np.float64)
        d, s = np.concatenate(ds), np.concatenate(ss)
        if self.final_add is not None and d.size:
            t = 1 - 1 / s
            add = self.final_add[d]
            dated = ~np.isnan(add)
            s = np.where(dated, (t + np.where(dated, add, 0.0)) / 10 ** 9, t)
        return d, s

    def search(self, q, limit=10):
        """``(top, total)`` with ``top`` = list of ``(score, docnum)`` in W11 order."""
        d, s = self.match_all(q)
        total = int(d.size)
        order = np.lexsort((d, -s))                                              # score desc, docnum asc
        if limit is not None:
            order = order[:limit]
        return [(float(s[i]), int(d[i])) for i in order], total
