#!/usr/bin/env python
"""Where does the time go?  Times query classes built from fixed rank bands (development tool).

Each class is 2000 queries of n terms drawn uniformly from a rank band of the config-2 corpus:
dense-only, sparse-only and mixed, OR and AND.  Prints postings/s and time per query.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from document_search_engine_b200 import And, Or, Term
    from document_search_engine_b200.corpus import config_corpus
    from document_search_engine_b200.scoring import BM25F
    from document_search_engine_b200.searching import Searcher
    variant = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    kw = {}
    for a in sys.argv[2:]:
        k, v = a.split("=")
        kw[k] = int(v)
    ix = config_corpus(2)
    s = Searcher(ix, weighting=BM25F, variant=variant, **kw)
    eng = s.engine
    rng = np.random.default_rng(1)
    bands = {"dense(51-150)": (51, 150), "mid(150-1500)": (150, 1500), "sparse(1500-20000)": (1500, 20000),
             "tiny(20000-200000)": (20000, 200000)}
    nq = 2000
    classes = []
    for name, (a, b) in bands.items():
        for mode in ("or", "and"):
            classes.append(("3x " + name + " " + mode, [(a, b)] * 3, mode))
    classes.append(("dense+2 sparse or", [(51, 150), (1500, 20000), (1500, 20000)], "or"))
    classes.append(("dense+2 sparse and", [(51, 150), (1500, 20000), (1500, 20000)], "and"))
    classes.append(("dense+2 tiny or", [(51, 150), (20000, 200000), (20000, 200000)], "or"))
    classes.append(("dense+2 tiny and", [(51, 150), (20000, 200000), (20000, 200000)], "and"))
    for name, bl, mode in classes:
        qs = []
        for _ in range(nq):
            terms = [Term("body", int(rng.integers(a, b))) for a, b in bl]
            qs.append(Or(terms) if mode == "or" else And(terms))
        batch = s.pack(qs)
        plan = eng.prepare(batch, 10)
        for _ in range(2):
            plan.execute()
        eng.synchronize()
        eng.reset_stats()
        for _ in range(3):
            plan.execute()
        eng.synchronize()
        st = eng.stats()
        ms = st["ms_score"] / max(1, st["n_executes"])
        P = st["postings_touched"]
        print(json.dumps({"class": name, "ms": round(ms, 3), "us_per_query": round(ms * 1e3 / nq, 3),
                          "postings_per_query": P // nq, "Gpostings_s": round(P / ms / 1e6, 1),
                          "algo_GBs": round(9 * P / ms / 1e6, 1)}))
        plan.close()


if __name__ == "__main__":
    main()
