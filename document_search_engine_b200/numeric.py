"""Field-length quantisation used by the scorable TEXT fields.

The reference declares its body/heading fields as Whoosh ``TEXT`` fields
(reference ``my_index.py:172-177``).  Whoosh 2.7.4 (``requirements.txt:6``,
not vendored) stores one *length byte* per (document, scorable field) and BM25F
decodes it back to an approximate length.  SURVEY.md W5/W6 restate the two
functions; this module is the host-side implementation the engine uses when it
builds the per-field norm tables that are uploaded to the GPU.
"""
from __future__ import annotations

from math import log

import numpy as np

#: lengths at or above this value all quantise to byte 255 (W6)
LENGTH_CLAMP = 108116


def length_to_byte(length) -> int:
    """Logarithmic 8-bit approximation of a field length (W6)."""
    if length is None:
        return 0
    if length >= LENGTH_CLAMP:
        return 255
    # Python 3 ``round`` (banker's rounding) on purpose: the table must be
    # generated the way the Python reference library generates it.
    return int(round(log((length / 27.0) + 1, 1.033)))


def byte_to_length(b: int) -> int:
    """Inverse table entry for a length byte (W6)."""
    return int(round((pow(1.033, b) - 1) * 27))


#: 256-entry decode table, int32 (strictly increasing from index 1)
B2L = np.array([byte_to_length(i) for i in range(256)], dtype=np.int32)


def lengths_to_bytes(lengths: np.ndarray) -> np.ndarray:
    """Vectorised ``length_to_byte`` that is bit-identical to the scalar one.

    The scalar function is evaluated once for every length up to the largest
    value present (bounded by ``LENGTH_CLAMP``), then gathered, so there is no
    second implementation of the rounding rule to drift.
    """
    lengths = np.asarray(lengths)
    if lengths.size == 0:
        return np.zeros(0, dtype=np.uint8)
    top = int(min(int(lengths.max()), LENGTH_CLAMP))
    table = _l2b_table(top)
    return table[np.minimum(lengths, top).astype(np.int64)]


_L2B_CACHE = np.zeros(0, dtype=np.uint8)


def _l2b_table(top: int) -> np.ndarray:
    global _L2B_CACHE
    if _L2B_CACHE.size <= top:
        n = max(top + 1, 8192)
        _L2B_CACHE = np.array([length_to_byte(i) for i in range(n)], dtype=np.uint8)
    return _L2B_CACHE


def doc_field_length_from_byte(b: int, default: int = 1) -> int:
    """W5: a zero length byte means "no length stored" and scores with ``default``."""
    return int(B2L[b]) if b else default


def norm_table(B: float, K1: float, avgfl: float) -> np.ndarray:
    """256-entry table ``K1 * ((1 - B) + B * fl / avgfl)`` in float64.

    ``fl`` is the decoded length for the byte, or 1 for byte 0 (W5).  The GPU
    consumes this table rounded once to float32; the oracle evaluates W1
    directly and never reads this table.
    """
    fl = B2L.astype(np.float64)
    fl[0] = 1.0
    return K1 * ((1.0 - B) + B * fl / float(avgfl))
