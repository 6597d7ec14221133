"""UK/US spelling variants.

The reference loads ``uk_us_variations.txt`` (154 "uk us" pairs, one per line) into two dicts
at process start (``my_flask.py:531-537``, ``:544-546``) and uses them in its "Did you mean"
path to swap a token for its other-dialect spelling when that spelling occurs in the index
(``my_flask.py:253-256``).  BASELINE config 3 defines a query rewrite on top of the same table:
every leaf ``Term(f, t)`` whose text has a variant becomes ``Or([Term(f, t), Term(f, other)])``,
so a 4-term AND becomes an AND of four 2-way ORs — the "AND of OR-groups" form the kernels
score natively.

The table itself is data of the reference deployment; it is read from the application directory
at run time (same relative path the reference opens) and is not copied into this repository.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, Optional, Set, Tuple

from .query import And, Or, Query, Term, _Compound

DEFAULT_FILE = "uk_us_variations.txt"


class Variants:
    def __init__(self, uk: Optional[Dict[str, str]] = None, us: Optional[Dict[str, str]] = None):
        self.uk_variations: Dict[str, str] = dict(uk or {})      # uk spelling -> us spelling
        self.us_variations: Dict[str, str] = dict(us or {})      # us spelling -> uk spelling
        self.uk_us_variations: Set[str] = set(self.uk_variations) | set(self.us_variations)

    @classmethod
    def load(cls, path: str = DEFAULT_FILE) -> "Variants":
        """Same parsing as the reference loader: ``uk, us = line.strip().split(' ')``."""
        v = cls()
        with open(path, encoding="utf-8", mode="r") as f:
            for line in f.readlines():
                if not line.strip():
                    continue
                uk, us = line.strip().split(" ")
                v.uk_variations[uk] = us
                v.us_variations[us] = uk
                v.uk_us_variations.add(uk)
                v.uk_us_variations.add(us)
        return v

    def other(self, word: str) -> Optional[str]:
        """The other-dialect spelling of ``word`` or ``None`` (uk table first, as the reference
        iterates ``(uk_variations, us_variations)``)."""
        for table in (self.uk_variations, self.us_variations):
            if word in table:
                return table[word]
        return None

    def substitute(self, word: str, frequency: Callable[[str], float]) -> str:
        """The reference's own use (``my_flask.py:253-256``): replace ``word`` by its variant if the
        variant occurs in the index (``frequency(variant) > 0``)."""
        o = self.other(word)
        return o if o is not None and frequency(o) > 0 else word

    def expand(self, q: Query, only_if: Optional[Callable[[str, str], bool]] = None) -> Query:
        """OR-expansion of every leaf that has a variant.  ``only_if(fieldname, variant)`` can veto a
        variant (e.g. ``lambda f, w: reader.frequency(f, w) > 0`` mirrors the reference's check)."""
        if isinstance(q, Term):
            o = self.other(q.text) if isinstance(q.text, str) else None
            if o is None or (only_if is not None and not only_if(q.fieldname, o)):
                return q
            return Or([Term(q.fieldname, q.text, boost=q.boost), Term(q.fieldname, o, boost=q.boost)])
        if isinstance(q, _Compound):
            return type(q)([self.expand(s, only_if) for s in q.subqueries], boost=q.boost)
        return q


def expand_with_map(q: Query, partner: Callable[[object], Optional[object]]) -> Query:
    """Same rewrite with an arbitrary variant map (the synthetic rank involution of config 3)."""
    if isinstance(q, Term):
        o = partner(q.text)
        if o is None:
            return q
        return Or([Term(q.fieldname, q.text, boost=q.boost), Term(q.fieldname, o, boost=q.boost)])
    if isinstance(q, _Compound):
        return type(q)([expand_with_map(s, partner) for s in q.subqueries], boost=q.boost)
    return q
