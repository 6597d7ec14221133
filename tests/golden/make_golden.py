#!/usr/bin/env python
"""Regenerates tests/golden/*.json.

PARITY UNPINNED: the reference's scoring path is Whoosh 2.7.4 (reference requirements.txt:6), which
is neither vendored under /root/reference nor installable in this image, and the reference holds
no tests or golden vectors of its own (SURVEY.md §4, §8 c).  These fixtures are therefore produced
by the doc-at-a-time restatement of Whoosh's semantics (oracle/whoosh_port.py, float64), plus the
hand-derived KAT-1 vector of SURVEY.md §8 c.  They pin the oracle against regressions and give the
GPU tests committed vectors to compare with; they are not outputs of Whoosh.

usage: python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def dump_queries(qs):
    out = []
    for q in qs:
        out.append(repr_query(q))
    return out


def repr_query(q):
    from document_search_engine_b200.query import And, Or, Term
    if isinstance(q, Term):
        return {"t": [q.fieldname, q.text, q.boost]}
    kind = "and" if isinstance(q, And) else "or"
    return {kind: [repr_query(s) for s in q.subqueries]}


def main():
    import numpy as np
    from document_search_engine_b200.corpus import make_corpus, make_queries, TITLE, BODY
    from oracle.whoosh_port import OracleSearcher
    cases = []
    # 1: config-1-shaped slice: one field, mixed 1-4-term AND/OR, densest terms included
    spec = dict(name="mixed_body", n_docs=3000, vocab=4000, seed=20260001, fields=["body"], qseed=20261001,
                n_queries=120, min_terms=1, max_terms=4, mode="mixed", skip_top=0, k=10, B=0.75, K1=1.2, field_B={})
    cases.append(spec)
    # 2: AND of OR-groups (config 3 shape: every term OR its variant partner)
    cases.append(dict(name="variants_groups", n_docs=3000, vocab=4000, seed=20260001, fields=["body"], qseed=20261003,
                      n_queries=60, min_terms=4, max_terms=4, mode="and", skip_top=0, k=10, variants=True, B=0.75, K1=1.2,
                      field_B={}))
    # 3: two fields, leaf boosts, per-field B (config 5 shape)
    cases.append(dict(name="two_fields", n_docs=2000, vocab=1500, seed=123, fields=["title", "body"], qseed=321,
                      n_queries=80, min_terms=1, max_terms=3, mode="mixed", skip_top=0, k=10, field_boosts=[2.0, 1.0],
                      B=0.6, K1=1.5, field_B={"title": 0.2}))
    # 4: deleted documents, zero length bytes, all weights 1 (exact score ties)
    cases.append(dict(name="deleted_ties", n_docs=2500, vocab=300, seed=99, fields=["body"], qseed=11, n_queries=80,
                      min_terms=1, max_terms=4, mode="mixed", skip_top=0, k=10, deleted_frac=0.2, zero_len_frac=0.1,
                      unit_tf=True, B=0.75, K1=1.2, field_B={}))
    # 5: NOT clauses (Whoosh's AndNotMatcher): every second AND / OR query excludes one or two dense terms
    cases.append(dict(name="not_clauses", n_docs=3000, vocab=4000, seed=20260001, fields=["body"], qseed=20261007,
                      n_queries=100, min_terms=2, max_terms=4, mode="mixed", skip_top=0, k=10, nots=True, B=0.75, K1=1.2,
                      field_B={}))
    for c in cases:
        ix, queries = build_case(c)
        o = OracleSearcher(ix, B=c["B"], K1=c["K1"], field_B=c["field_B"])
        results = []
        for q in queries:
            top, total = o.search(q, limit=c["k"])
            results.append({"total": int(total), "top": [[float(s), int(d)] for s, d in top]})
        fixture = {"_about": "oracle/whoosh_port.py output (float64); NOT a Whoosh run - parity unpinned, see make_golden.py",
                   "case": c, "results": results}
        with open(os.path.join(HERE, c["name"] + ".json"), "w") as f:
            json.dump(fixture, f, separators=(",", ":"))
        print(c["name"], len(results), "queries")


def build_case(c):
    """Corpus and queries of a fixture, regenerated from its recorded parameters."""
    import numpy as np
    from document_search_engine_b200.corpus import make_corpus, make_queries, TITLE, BODY
    specs = (TITLE, BODY) if len(c["fields"]) == 2 else (BODY,)
    ix = make_corpus(c["n_docs"], c["vocab"], c["seed"], specs, device="cpu")
    if c.get("deleted_frac"):
        rng = np.random.default_rng(0)
        ix.deleted = (rng.random(ix.n_docs_all) < c["deleted_frac"]).astype(np.uint8)
        ix.len_bytes[0, rng.random(ix.n_docs_all) < c["zero_len_frac"]] = 0
    if c.get("unit_tf"):
        ix.tfs[:] = 1.0
    qs = make_queries(c["n_queries"], c["vocab"], c["qseed"], c["min_terms"], c["max_terms"], c["mode"],
                      fields=tuple(c["fields"]), field_boosts=tuple(c.get("field_boosts", [1.0] * len(c["fields"]))),
                      variants=c.get("variants", False), skip_top=c["skip_top"])
    queries = qs.queries
    if c.get("nots"):
        from document_search_engine_b200.query import And, Not, Or, Term
        out = []
        for i, q in enumerate(queries):
            if i % 2 == 0 and isinstance(q, (And, Or)):
                neg = [Term(c["fields"][0], (i * 7) % 150)] + ([Term(c["fields"][0], (i * 13) % 400)] if i % 4 == 0 else [])
                q = type(q)(list(q.subqueries) + [Not(neg[0] if len(neg) == 1 else Or(neg))])
            out.append(q)
        queries = out
    return ix, queries


if __name__ == "__main__":
    main()
