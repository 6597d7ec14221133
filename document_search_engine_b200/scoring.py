"""Weighting models with the ``whoosh.scoring`` surface the reference uses.

The reference passes the *class* ``BM25F`` (or a subclass with a ``final`` hook)
to ``ix.searcher(weighting=...)`` (``my_flask.py:183-184``) and subclasses it in
``my_whoosh.py:127-154``.  W1-W4 of SURVEY.md §8 c give the arithmetic; the
host side evaluates everything that is per-term or per-field in float64 and
hands the GPU two float32 products:

* per leaf   ``w = idf * (K1 + 1) * boost``
* per field  ``norm[b] = K1 * ((1 - B_field) + B_field * fl(b) / avgfl)``

so that the per-posting work on the device is ``w * tf / (tf + norm[lb])``.
"""
from __future__ import annotations

import re
from datetime import datetime, timedelta
from math import log

import numpy as np

from .numeric import norm_table


class WeightingModel:
    use_final = False

    def idf(self, searcher, fieldname, text) -> float:
        """W3: ``log(dc / (df + 1)) + 1`` with corpus-wide dc (incl. deleted) and stored df."""
        n = searcher.doc_frequency(fieldname, text)
        dc = searcher.doc_count_all()
        return log(dc / (n + 1)) + 1

    def final(self, searcher, docnum, score):
        return score


class BM25F(WeightingModel):
    """``BM25F(B=0.75, K1=1.2, **{"<field>_B": b})`` (W2)."""

    def __init__(self, B=0.75, K1=1.2, **kwargs):
        self.B = float(B)
        self.K1 = float(K1)
        self._field_B = {}
        for k, v in kwargs.items():
            if k.endswith("_B"):
                self._field_B[k[:-2]] = float(v)
            else:
                raise TypeError("unexpected BM25F argument %r" % k)

    def field_B(self, fieldname) -> float:
        return self._field_B.get(fieldname, self.B)

    def supports_block_quality(self):
        return True

    # -- host-side products consumed by the engine ---------------------------
    def leaf_weight(self, searcher, fieldname, text, boost=1.0) -> float:
        """W15: Whoosh's ``BM25F.scorer`` returns a ``WeightScorer`` for a field that is not scorable (the
        reference's ``book=ID`` field, ``my_index.py:152``; the UI's book filter generates ``book:xyz`` terms,
        ``static/main.js:5-16``): the score is the posting weight, times the boost."""
        if not searcher.stats_ix.is_scorable(fieldname):
            return float(boost)
        return self.idf(searcher, fieldname, text) * (self.K1 + 1.0) * boost

    def norm_tables(self, ix) -> np.ndarray:
        """float32 ``[n_fields, 256]`` norm tables for index ``ix`` (global avgfl, W4/W8)."""
        out = np.empty((len(ix.field_names), 256), dtype=np.float32)
        for f, name in enumerate(ix.field_names):
            if not ix.is_scorable(name):
                out[f] = -1.0               # W15: WeightScorer, the posting weight is the score
            else:
                out[f] = norm_table(self.field_B(name), self.K1, ix.avg_field_length(name)).astype(np.float32)
        return out

    def norm_key(self):
        """What the per-posting impacts depend on (not the class: ``DescDateBM25F`` has ``BM25F``'s norms)."""
        return (self.B, self.K1, tuple(sorted(self._field_B.items())))

    def key(self):
        return ("BM25F",) + self.norm_key()


class DateBM25F(BM25F):
    """BM25F whose ``final`` step orders hits by date first (reference ``my_whoosh.py:127-154``, chosen at
    ``my_flask.py:183`` for two of the UI's three hit orders).  Whoosh calls ``final`` for every match
    before the top-k collector (W14), so the device has to apply it: the per-document part (the date, plus
    the chapter number found in the stored heading, as seconds) is computed once on the host by
    ``doc_final_terms`` and uploaded; the kernels evaluate, in float64,

        v = 1 - 1/score                           (document without a date)
        v = ((1 - 1/score) + (date_s + 1.0)) / 10**9   (dated document)

    ``final`` below is the same arithmetic for one document (what the oracle calls)."""
    use_final = True
    #: seconds are counted up from this instant (newest first) or down to it (oldest first)
    descending = True
    _EPOCH_DESC = datetime(1800, 1, 1)
    _EPOCH_ASC = datetime(2200, 1, 1)
    _CHAPTER = re.compile(r"chapter\W*(\d+)", re.IGNORECASE)

    def date_seconds(self, fields):
        """Date score of a document's stored fields, or ``None`` when it has no date."""
        if "date" not in fields:
            return None
        m = self._CHAPTER.search(fields.get("heading") or "")
        when = fields["date"] + timedelta(seconds=int(m.group(1)) if m else 0)
        if self.descending:
            return (when - self._EPOCH_DESC).total_seconds()
        return (self._EPOCH_ASC - when).total_seconds()

    def final(self, searcher, docnum, score):
        v = 1 - 1 / score
        ds = self.date_seconds(searcher.stored_fields(docnum))
        if ds is not None:
            v += ds + 1.0
            v /= 10 ** 9
        return v

    def doc_final_terms(self, ix) -> np.ndarray:
        """float64 ``[n_docs_all]``: ``date_s + 1.0`` per document of ``ix``, NaN where there is no date."""
        out = np.full(ix.n_docs_all, np.nan, dtype=np.float64)
        if ix.stored is not None:
            for i, fields in enumerate(ix.stored):
                ds = self.date_seconds(fields)
                if ds is not None:
                    out[i] = ds + 1.0
        return out

    def key(self):
        return (type(self).__name__,) + super().key()[1:]


class DescDateBM25F(DateBM25F):
    descending = True


class AscDateBM25F(DateBM25F):
    descending = False


def bm25(idf, tf, fl, avgfl, B, K1):
    """W1, in exactly this association (float64 when given Python floats)."""
    return idf * ((tf * (K1 + 1)) / (tf + K1 * ((1 - B) + B * fl / avgfl)))


def instantiate(weighting) -> WeightingModel:
    """A weighting *class* is instantiated with defaults (W2; ``my_flask.py:183``)."""
    if weighting is None:
        return BM25F()
    if isinstance(weighting, type):
        return weighting()
    return weighting
