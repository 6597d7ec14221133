"""Whoosh index -> ``FlatIndex``.

The reference opens its index with ``whoosh.index.open_dir`` (``my_index.py:226-234``) over the schema of
``my_index.py:163-178`` and hands ``ix`` to the front ends (``my_flask.py:549``, ``cli.py:25``).  ``flatten_reader``
turns what a Whoosh ``IndexReader`` exposes through its PUBLIC API into the arrays the engine uploads:

    reader.indexed_field_names() / reader.schema[name]     which fields, which of them are scorable (W15)
    reader.lexicon(name)                                    the field's terms, sorted (bytes)
    reader.doc_frequency(name, btext)                       df AS STORED: deleted documents still counted (W3);
                                                            the reference deletes and re-adds every document
                                                            (``my_index.py:115-117``), so this is not the live count
    reader.postings(name, btext)                            matcher: is_active / id / weight / next
    reader.doc_field_length(docnum, name, 0)                quantised by length_to_byte (W5, W6)
    reader.field_length(name)                               exact token total (W4)
    reader.doc_count_all() / reader.is_deleted(docnum)      W3, W9
    reader.stored_fields(docnum)                            ``hit['session']`` etc. (``my_flask.py:315``, ``my_whoosh.py:131``)

Whoosh itself is not a dependency of this module: anything with those methods works, which is how it is tested
here (``tests/test_flatten.py`` drives it with a stand-in reader over a ``FlatIndex``; ``tests/test_whoosh_pin.py``
with a real index when Whoosh is installed).  A multi-segment index is read through its top-level reader, whose
docnums are already global (W8).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import numpy as np

from .index import FlatIndex
from .numeric import lengths_to_bytes


def _text(field, btext):
    """Term bytes -> the text the query side uses (``field.from_bytes`` when the schema type has it)."""
    fb = getattr(field, "from_bytes", None)
    if fb is not None:
        return fb(btext)
    return btext.decode("utf-8") if isinstance(btext, (bytes, bytearray)) else btext


def flatten_reader(reader, fields: Optional[Sequence[str]] = None, stored: bool = True,
                   date_fields: Sequence[str] = ()) -> FlatIndex:
    """Flatten the postings of ``fields`` (default: every indexed field) of a Whoosh-style ``reader``.
    ``date_fields`` (the reference's ``date``): instead of Whoosh's tiered numeric terms these fields get the flat
    index's own year / month / day tokens, built from the documents' stored values (``dates.py``), which is what
    ``DateRange`` queries expand to."""
    schema = reader.schema
    names = list(fields) if fields is not None else list(reader.indexed_field_names())
    names = [n for n in names if n not in date_fields]
    n_docs = int(reader.doc_count_all())
    offs: List[int] = [0]
    docids: List[np.ndarray] = []
    tfs: List[np.ndarray] = []
    term_field: List[int] = []
    df: List[int] = []
    terms = {}
    scorable = []
    pos_triples = {}                      # field -> (docs, positions, term ids) of every occurrence, for phrases
    for f, name in enumerate(names):
        field = schema[name]
        scorable.append(bool(getattr(field, "scorable", False)))
        for btext in reader.lexicon(name):
            m = reader.postings(name, btext)
            has_pos = callable(getattr(m, "supports", None)) and m.supports("positions")
            ids, ws = [], []
            # deleted documents stay in the flat postings with the ``deleted`` flags beside them: the library drops
            # them at upload (W9), and a raw matcher may or may not have filtered them already
            while m.is_active():
                ids.append(m.id())
                ws.append(m.weight())
                if has_pos:
                    # Whoosh's positions format (TEXT(phrase=True), reference my_index.py:172-177): the word order
                    ps = list(m.value_as("positions"))
                    t3 = pos_triples.setdefault(f, ([], [], []))
                    t3[0].extend([ids[-1]] * len(ps))
                    t3[1].extend(ps)
                    t3[2].extend([len(term_field)] * len(ps))
                m.next()
            a = np.asarray(ids, dtype=np.uint32)
            if a.size > 1 and not (a[1:] > a[:-1]).all():
                order = np.argsort(a, kind="stable")
                a, ws = a[order], [ws[i] for i in order]
            terms[(f, _text(field, btext))] = len(term_field)
            term_field.append(f)
            df.append(int(reader.doc_frequency(name, btext)))
            docids.append(a)
            tfs.append(np.asarray(ws, dtype=np.float32))
            offs.append(offs[-1] + a.size)
    lengths = np.zeros((len(names), n_docs), dtype=np.int64)
    totals = np.zeros(len(names), dtype=np.uint64)
    for f, name in enumerate(names):
        totals[f] = int(reader.field_length(name))
        if scorable[f]:
            for d in range(n_docs):
                lengths[f, d] = reader.doc_field_length(d, name, 0)
    deleted = np.fromiter((1 if reader.is_deleted(d) else 0 for d in range(n_docs)), dtype=np.uint8, count=n_docs)
    stored_docs = None
    if stored or date_fields:
        stored_docs = [dict(reader.stored_fields(d)) if not deleted[d] else {} for d in range(n_docs)]
    if date_fields:
        from .dates import tier_tokens
        for name in date_fields:
            f = len(names)
            names.append(name)
            scorable.append(False)
            lists = {}
            for d, sf in enumerate(stored_docs):
                if sf.get(name) is not None:
                    for tok in tier_tokens(sf[name]):
                        lists.setdefault(tok, []).append(d)
            for tok in sorted(lists):
                a = np.asarray(lists[tok], dtype=np.uint32)
                terms[(f, tok)] = len(term_field)
                term_field.append(f)
                df.append(int(a.size))
                docids.append(a)
                tfs.append(np.ones(a.size, dtype=np.float32))
                offs.append(offs[-1] + a.size)
        lengths = np.concatenate([lengths, np.zeros((len(date_fields), n_docs), dtype=np.int64)])
        totals = np.concatenate([totals, np.zeros(len(date_fields), dtype=np.uint64)])
        if not stored:
            stored_docs = None
    positions = {}
    for f, (pd, pp, pt) in pos_triples.items():
        pd, pp, pt = np.asarray(pd, np.int64), np.asarray(pp, np.int64), np.asarray(pt, np.int32)
        dlen = np.zeros(n_docs, dtype=np.int64)
        np.maximum.at(dlen, pd, pp + 1)                       # a gap (a removed stop word) stays -1
        po = np.zeros(n_docs + 1, dtype=np.int64)
        np.cumsum(dlen, out=po[1:])
        seq = np.full(int(po[-1]), -1, dtype=np.int32)
        seq[po[pd] + pp] = pt
        positions[f] = (po, seq)
    cat = (lambda parts, dt: np.concatenate(parts).astype(dt) if parts else np.zeros(0, dt))
    return FlatIndex(positions=positions, field_names=names, n_docs_all=n_docs, term_offsets=np.asarray(offs, dtype=np.uint64),
                     docids=cat(docids, np.uint32), tfs=cat(tfs, np.float32),
                     term_field=np.asarray(term_field, dtype=np.uint8),
                     len_bytes=np.stack([lengths_to_bytes(lengths[f]) for f in range(len(names))]) if names else np.zeros((0, n_docs), np.uint8),
                     field_length_total=totals, terms=terms, df=np.asarray(df, dtype=np.uint32),
                     deleted=deleted if deleted.any() else None, stored=stored_docs, scorable=scorable)


def flatten_index(ix, fields: Optional[Sequence[str]] = None, stored: bool = True, date_fields: Sequence[str] = ()) -> FlatIndex:
    """``flatten_reader`` over ``ix.reader()`` (a ``whoosh.index.Index``, e.g. ``my_index.get_idx('index')``)."""
    reader = ix.reader()
    try:
        return flatten_reader(reader, fields=fields, stored=stored, date_fields=date_fields)
    finally:
        close = getattr(reader, "close", None)
        if close is not None:
            close()
