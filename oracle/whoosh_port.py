"""CPU ORACLE (test infrastructure, not product code) — doc-at-a-time port.

PARITY UNPINNED: the algorithm on this path lives in the third-party library
Whoosh 2.7.4 (reference ``requirements.txt:6``), which is neither vendored under
``/root/reference`` nor installable here (no network, no wheel).  The reference
holds no tests, golden vectors or fixtures for this path (SURVEY.md §4, §8 c).
This file therefore restates the *published* Whoosh 2.7.4 behaviour, item by
item as W1-W14 of SURVEY.md §8 c, and is anchored on the reference's call sites:

* weighting selection / defaults ........ ``my_flask.py:183-184``  (W2)
* ``search_page`` / ``search`` limits ..... ``my_flask.py:208``, ``:211``, ``:304``, ``cli.py:9``
* ``final`` hook contract ................ ``my_whoosh.py:127-154``  (W14)
* which fields are scorable TEXT ........ ``my_index.py:172-177``; ``book=ID`` is not: ``:152``, ``:171``  (W15)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this module.  It deliberately mirrors the shape of the library it
restates (one Python-level step per matching posting: matcher tree → scorer →
heap collector), because that is what the CPU baseline is meant to time.

The index argument is duck-typed: any object with ``term_offsets``, ``docids``,
``tfs``, ``len_bytes[f, d]``, ``field_length_total``, ``df``, ``deleted``,
``doc_base``, ``field_names``, ``doc_count_all()`` and ``term_id(field, text)``.
Queries are duck-typed on the class names ``Term`` / ``And`` / ``Or`` / ``Not`` / ``Every``.
"""
from __future__ import annotations

from bisect import bisect_left
from heapq import heappush, heapreplace
from math import log

# ---------------------------------------------------------------------------
# W6 / W5: length quantisation (Whoosh util/numeric.py)
# ---------------------------------------------------------------------------


def length_to_byte(length):
    if length is None:
        return 0
    if length >= 108116:
        return 255
    return int(round(log((length / 27.0) + 1, 1.033)))


_B2L = [int(round((pow(1.033, i) - 1) * 27)) for i in range(256)]


def byte_to_length(b):
    return _B2L[b]


# ---------------------------------------------------------------------------
# W1-W4: BM25F (Whoosh scoring.py)
# ---------------------------------------------------------------------------


def bm25(idf, tf, fl, avgfl, B, K1):
    # W1, in exactly this association
    return idf * ((tf * (K1 + 1)) / (tf + K1 * ((1 - B) + B * fl / avgfl)))


class OracleSearcher:
    """Top-level ("parent") searcher: owns the corpus-wide statistics (W3, W4, W8)."""

    def __init__(self, ix, B=0.75, K1=1.2, field_B=None, final=None, shards=None):
        self.ix = ix
        self.B = B
        self.K1 = K1
        self.field_B = dict(field_B or {})
        self.final = final              # W14: callable(searcher, docnum, score) or None
        # W8: a multi-segment index is a list of sub-indexes with doc offsets
        self.shards = list(shards) if shards is not None else [ix]
        self._idf = {}

    def doc_count_all(self):
        return self.ix.doc_count_all()

    def doc_frequency(self, fieldname, text):
        tid = self.ix.term_id(fieldname, text)
        return 0 if tid < 0 else int(self.ix.df[tid])

    def idf(self, fieldname, text):
        key = (fieldname, text)
        v = self._idf.get(key)
        if v is None:
            # W3: natural log, true division, dc includes deleted documents
            v = log(self.doc_count_all() / (self.doc_frequency(fieldname, text) + 1)) + 1
            self._idf[key] = v
        return v

    def avg_field_length(self, fieldname):
        f = self.ix.field_names.index(fieldname)
        # W4
        return (int(self.ix.field_length_total[f]) / (self.doc_count_all() or 1)) or 1

    # -- searching ----------------------------------------------------------
    def search(self, q, limit=10):
        """Returns ``(top, total)``: ``top`` is a list of ``(score, global docnum)`` in
        W11 order; ``total`` is the exact number of matching documents (W13)."""
        heap = []
        allhits = []
        total = 0
        for sub in self.shards:
            m = self._matcher(q, sub)
            base = sub.doc_base
            while m.is_active():
                d = m.id() + base
                score = m.score()
                if self.final is not None:
                    score = self.final(self, d, score)          # W14: before the heap
                total += 1
                if limit is None:
                    allhits.append((score, -d))                 # UnlimitedCollector
                elif len(heap) < limit:
                    heappush(heap, (score, -d))
                elif score > heap[0][0]:                        # W11: strict >
                    heapreplace(heap, (score, -d))
                m.next()
        items = allhits if limit is None else heap
        items.sort(reverse=True)                                # score desc, docnum asc
        return [(s, -nd) for s, nd in items], total

    # -- key terms / more-like-this (Whoosh classify.py Expander + Bo1Model, searching.py more_like) ------------
    def key_terms(self, vector, fieldname, numterms=5, normalize=True):
        """``vector``: iterable of (word, weight) - the text's tokens with weight 1, or a document's term vector."""
        ix = self.ix
        N = float(ix.doc_count_all())
        top = {}
        for word, weight in vector:
            top[word] = top.get(word, 0) + weight
        tlist, maxweight = [], 0
        for word, weight in top.items():
            tid = ix.term_id(fieldname, word)
            if tid < 0:
                continue
            a, b = int(ix.term_offsets[tid]), int(ix.term_offsets[tid + 1])
            cf = float(sum(ix.tfs[a:b].tolist()))                  # reader.frequency: total weight in the collection
            f = cf / N
            score = weight * log((1.0 + f) / f, 2) + log(1.0 + f, 2)
            if score > maxweight:
                maxweight = score
            tlist.append((score, word))
        if not tlist:
            return []
        if normalize:
            f = maxweight / N
            norm = (maxweight * log((1.0 + f) / f) + log(1.0 + f)) / log(2.0)
        else:
            norm = maxweight
        tlist = [(weight / norm, t) for weight, t in tlist]
        tlist.sort(key=lambda x: (0 - x[0], x[1]))
        return [(t, weight) for weight, t in tlist[:numterms]]

    def more_like(self, docnum, fieldname, vector, top=10, numterms=5):
        """``Or`` of the key terms with their weights as boosts, document ``docnum`` masked (never collected)."""
        kts = self.key_terms(vector, fieldname, numterms=numterms)
        if not kts:
            return [], 0
        q = _O([_T(fieldname, w, boost=wt) for w, wt in kts])
        hits, total = self.search(q, limit=None)
        keep = [(s, d) for s, d in hits if d != docnum]
        return keep[:top], len(keep)

    def _matcher(self, q, sub):
        name = type(q).__name__
        if name == "Term":
            tid = sub.term_id(q.fieldname, q.text)
            if tid < 0:
                return NullMatcher()                            # W10: unknown term/field
            a, b = int(sub.term_offsets[tid]), int(sub.term_offsets[tid + 1])
            f = sub.field_names.index(q.fieldname)
            if not _scorable(sub, f):
                # W15: BM25F.scorer() returns WeightScorer for a field whose schema type is not scorable
                # (reference book=ID, my_index.py:152, :171; queried by the UI's book filter, static/main.js:5-16)
                scorer = WeightScorer()
            else:
                scorer = BM25FScorer(self, q.fieldname, q.text,
                                     self.field_B.get(q.fieldname, self.B), self.K1)
            m = PostingMatcher(sub.docids[a:b].tolist(), sub.tfs[a:b].tolist(),
                               sub.len_bytes[f], scorer, sub.deleted)
            return m if q.boost == 1.0 else BoostMatcher(m, q.boost)
        if name == "Every":
            f = sub.field_names.index(q.fieldname) if q.fieldname in sub.field_names else -1
            if f < 0:
                return NullMatcher()
            lb = sub.len_bytes[f]
            dele = sub.deleted
            ids = [d for d in range(lb.shape[0]) if lb[d] and not (dele is not None and dele[d])]
            return ConstMatcher(ids, q.boost)
        if name in ("And", "Or"):
            # Whoosh's compound matcher takes the Not children out first, builds the matcher of the
            # rest, and wraps it in AndNotMatcher(rest, union of the negated queries).
            nots = [s.query for s in q.subqueries if type(s).__name__ == "Not"]
            subs = [self._matcher(s, sub) for s in q.subqueries if type(s).__name__ != "Not"]
            if not subs:
                return NullMatcher()
            cls = IntersectionMatcher if name == "And" else UnionMatcher
            m = _binary_tree(cls, subs)
            if nots:
                notm = _binary_tree(UnionMatcher, [self._matcher(s, sub) for s in nots])
                if notm.is_active():
                    m = AndNotMatcher(m, notm)
            return m if q.boost == 1.0 else BoostMatcher(m, q.boost)
        if name in ("Prefix", "Wildcard"):
            # Whoosh query/terms.py MultiTerm.matcher: the words of the field's lexicon that fit the pattern, in
            # lexicon order; none -> NullMatcher, one -> that term's matcher (without the pattern's boost), else the
            # Or of the terms with the pattern's boost.  Scored by the searcher's weighting (Searcher.postings falls
            # back to it although pattern queries ask for a constant score).
            f = sub.field_names.index(q.fieldname) if q.fieldname in sub.field_names else -1
            words = sorted(t for (ff, t) in (sub.terms or {}) if ff == f and isinstance(t, str) and _pattern_fits(q, t))
            ms = [self._matcher(_T(q.fieldname, w), sub) for w in words]
            if not ms:
                return NullMatcher()
            if len(ms) == 1:
                return ms[0]
            m = _binary_tree(UnionMatcher, ms)
            return m if q.boost == 1.0 else BoostMatcher(m, q.boost)
        if name == "Phrase":
            # Whoosh query.Phrase.matcher: a word missing from the field -> NullMatcher; else SpanNear.build(terms, slop,
            # ordered) = a left-deep chain of SpanNear2 over IntersectionMatchers: the positions decide whether a
            # document matches, its score is the intersection's (the sum of the words' scores); then the boost.
            if not q.words:
                return NullMatcher()
            ms = [self._matcher(_T(q.fieldname, w), sub) for w in q.words]
            if any(isinstance(m, NullMatcher) for m in ms):
                return NullMatcher()
            m = ms[0]
            for x in ms[1:]:
                m = IntersectionMatcher(m, x)
            if len(ms) > 1:
                from oracle.numpy_oracle import phrase_docs
                m = FilterMatcher(m, set(int(d) for d in phrase_docs(sub, q.fieldname, q.words, q.slop)))
            return m if q.boost == 1.0 else BoostMatcher(m, q.boost)
        if name == "DateRange":
            # Whoosh query.DateRange = NumericRange over the DATETIME field inside ConstantScoreQuery(boost): every
            # live document whose date lies in the range scores the boost (reference search-form.html:26, :39;
            # my_flask.py:189-193).  Evaluated over the stored dates, not through the index's date tokens.
            from oracle.numpy_oracle import date_range_docs
            ids = [int(d) for d in date_range_docs(sub, q.fieldname, q.start, q.end)
                   if not (sub.deleted is not None and sub.deleted[d])]
            return ConstMatcher(ids, q.boost) if ids else NullMatcher()
        if name == "_Null":
            return NullMatcher()
        raise NotImplementedError(name)


class _T:
    """A plain term (duck-typed like the engine's ``Term``) for expansions."""

    def __init__(self, fieldname, text, boost=1.0):
        self.fieldname, self.text, self.boost = fieldname, text, boost


_T.__name__ = "Term"


class _O:
    def __init__(self, subqueries, boost=1.0):
        self.subqueries, self.boost = subqueries, boost


_O.__name__ = "Or"


def _pattern_fits(q, word):
    if type(q).__name__ == "Prefix":
        return word.startswith(q.text)
    import fnmatch
    return fnmatch.fnmatchcase(word, q.text)


def _binary_tree(cls, ms):
    if len(ms) == 1:
        return ms[0]
    half = len(ms) // 2
    return cls(_binary_tree(cls, ms[:half]), _binary_tree(cls, ms[half:]))


def _scorable(ix, f):
    flags = getattr(ix, "scorable", None)
    return True if flags is None else bool(flags[f])


class WeightScorer:
    """W15 (Whoosh scoring.WeightScorer): the score of a posting is its weight; no idf, no length norm."""

    def score(self, weight, length_byte):
        return weight


class BM25FScorer:
    def __init__(self, parent, fieldname, text, B, K1):
        self.idf = parent.idf(fieldname, text)                  # parent searcher (W3, W8)
        self.avgfl = parent.avg_field_length(fieldname) or 1
        self.B = B
        self.K1 = K1

    def score(self, weight, length_byte):
        fl = byte_to_length(length_byte) if length_byte else 1  # W5
        return bm25(self.idf, weight, fl, self.avgfl, self.B, self.K1)


# ---------------------------------------------------------------------------
# Matchers (Whoosh matching/)
# ---------------------------------------------------------------------------


class NullMatcher:
    def is_active(self):
        return False

    def id(self):
        raise IndexError

    def next(self):
        pass

    def skip_to(self, d):
        pass

    def score(self):
        return 0.0


class PostingMatcher:
    """Leaf: one term's postings, deleted documents filtered out (W9)."""

    def __init__(self, ids, weights, len_bytes, scorer, deleted):
        self.ids = ids
        self.weights = weights
        self.lb = len_bytes
        self.scorer = scorer
        self.deleted = deleted
        self.i = 0
        self._skip_deleted()

    def _skip_deleted(self):
        if self.deleted is not None:
            ids, dele, n = self.ids, self.deleted, len(self.ids)
            while self.i < n and dele[ids[self.i]]:
                self.i += 1

    def is_active(self):
        return self.i < len(self.ids)

    def id(self):
        return self.ids[self.i]

    def next(self):
        self.i += 1
        self._skip_deleted()

    def skip_to(self, d):
        # Whoosh skips whole posting blocks and bisects inside one; a bisect from the
        # current position is the same O(log n) behaviour
        if self.i < len(self.ids) and self.ids[self.i] < d:
            self.i = bisect_left(self.ids, d, self.i + 1)
        self._skip_deleted()

    def score(self):
        return self.scorer.score(self.weights[self.i], int(self.lb[self.ids[self.i]]))


class ConstMatcher:
    """``Every``: every listed document scores the (boost) constant."""

    def __init__(self, ids, weight):
        self.ids = ids
        self.w = weight
        self.i = 0

    def is_active(self):
        return self.i < len(self.ids)

    def id(self):
        return self.ids[self.i]

    def next(self):
        self.i += 1

    def skip_to(self, d):
        while self.i < len(self.ids) and self.ids[self.i] < d:
            self.i += 1

    def score(self):
        return self.w


class FilterMatcher:
    """The span filter of a phrase: the child's documents that are in ``allowed``, scored by the child."""

    def __init__(self, child, allowed):
        self.c = child
        self.allowed = allowed
        self._find()

    def _find(self):
        while self.c.is_active() and self.c.id() not in self.allowed:
            self.c.next()

    def is_active(self):
        return self.c.is_active()

    def id(self):
        return self.c.id()

    def next(self):
        self.c.next()
        self._find()

    def skip_to(self, d):
        self.c.skip_to(d)
        self._find()

    def score(self):
        return self.c.score()


class BoostMatcher:
    def __init__(self, child, boost):
        self.child = child
        self.boost = boost

    def is_active(self):
        return self.child.is_active()

    def id(self):
        return self.child.id()

    def next(self):
        self.child.next()

    def skip_to(self, d):
        self.child.skip_to(d)

    def score(self):
        return self.child.score() * self.boost


class UnionMatcher:
    """W10: documents in either child; score = sum of the children positioned there."""

    def __init__(self, a, b):
        self.a = a
        self.b = b

    def is_active(self):
        return self.a.is_active() or self.b.is_active()

    def id(self):
        a, b = self.a, self.b
        if not a.is_active():
            return b.id()
        if not b.is_active():
            return a.id()
        return min(a.id(), b.id())

    def next(self):
        a, b = self.a, self.b
        aa, ba = a.is_active(), b.is_active()
        if aa and ba:
            ai, bi = a.id(), b.id()
            if ai <= bi:
                a.next()
            if bi <= ai:
                b.next()
        elif aa:
            a.next()
        elif ba:
            b.next()

    def skip_to(self, d):
        self.a.skip_to(d)
        self.b.skip_to(d)

    def score(self):
        a, b = self.a, self.b
        if not a.is_active():
            return b.score()
        if not b.is_active():
            return a.score()
        ai, bi = a.id(), b.id()
        if ai < bi:
            return a.score()
        if bi < ai:
            return b.score()
        return a.score() + b.score()


class IntersectionMatcher:
    """W10: documents in both children; score = sum of both."""

    def __init__(self, a, b):
        self.a = a
        self.b = b
        self._find()

    def _find(self):
        a, b = self.a, self.b
        while a.is_active() and b.is_active():
            ai, bi = a.id(), b.id()
            if ai == bi:
                return
            if ai < bi:
                a.skip_to(bi)
            else:
                b.skip_to(ai)

    def is_active(self):
        return self.a.is_active() and self.b.is_active()

    def id(self):
        return self.a.id()

    def next(self):
        self.a.next()
        self._find()

    def skip_to(self, d):
        self.a.skip_to(d)
        self.b.skip_to(d)
        self._find()

    def score(self):
        return self.a.score() + self.b.score()


class AndNotMatcher:
    """Whoosh matching.binary.AndNotMatcher: documents of ``a`` that are not in ``b``; score = a's."""

    def __init__(self, a, b):
        self.a = a
        self.b = b
        self._find()

    def _find(self):
        a, b = self.a, self.b
        while a.is_active():
            if b.is_active() and b.id() < a.id():
                b.skip_to(a.id())
            if b.is_active() and b.id() == a.id():
                a.next()
                continue
            return

    def is_active(self):
        return self.a.is_active()

    def id(self):
        return self.a.id()

    def next(self):
        self.a.next()
        self._find()

    def skip_to(self, d):
        self.a.skip_to(d)
        self._find()

    def score(self):
        return self.a.score()


# ---------------------------------------------------------------------------
# W13: search_page arithmetic (Whoosh searching.ResultsPage)
# ---------------------------------------------------------------------------


def page_view(total, pagenum, pagelen):
    """Returns ``(pagenum, offset, pagelen, pagecount)`` after Whoosh's clamping."""
    if pagenum < 1:
        raise ValueError("pagenum must be >= 1")
    pagecount = -(-total // pagelen)
    pagenum = min(pagecount, pagenum)
    offset = (pagenum - 1) * pagelen
    if offset + pagelen > total:
        pagelen = total - offset
    return pagenum, offset, pagelen, pagecount
