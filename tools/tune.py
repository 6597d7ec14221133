#!/usr/bin/env python
"""Time the hot path for several engine option sets on one GPU (development tool).

Generates the BASELINE config corpus once, then for every option set (``name=value,name=value`` with
``bm25f_options`` field names, ``default`` for the library defaults) creates an engine, checks a query sample
against the oracle, and times ``execute`` with the library's CUDA events.  One JSON line per set and query mix.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def parse_opts(text):
    if text in ("", "default"):
        return {}
    out = {}
    for part in text.split(","):
        name, _, value = part.partition("=")
        out[name.strip()] = int(value, 0)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--docs", type=int, default=0)
    ap.add_argument("--queries", type=int, default=0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--check", type=int, default=16)
    ap.add_argument("--opts", nargs="*", default=["default"])
    ap.add_argument("--modes", nargs="*", default=["cfg"], help="cfg | and | or : query mix to time")
    ap.add_argument("--arena", action="store_true", help="plan in the engine's workspaces (the path bm25f_search_batch / bm25f_submit take: device planner)")
    args = ap.parse_args()
    from document_search_engine_b200.corpus import CONFIGS, config_corpus, config_queries, make_queries
    from document_search_engine_b200.scoring import BM25F
    from document_search_engine_b200.searching import Searcher
    from oracle.numpy_oracle import NumpyOracle
    from tests.parity import assert_query_parity

    c = CONFIGS[args.config]
    t0 = time.time()
    ix = config_corpus(args.config, n_docs=args.docs or None)
    nq = args.queries or c["n_queries"]
    print("corpus %d docs %d postings in %.1fs" % (ix.n_docs_all, ix.n_postings, time.time() - t0), file=sys.stderr)
    o = NumpyOracle(ix)
    qsets = {}
    for m in args.modes:
        if m == "cfg":
            qsets[m] = config_queries(args.config, nq).queries
        else:
            qsets[m] = make_queries(nq, c["vocab"], 20261000 + args.config, c["min_terms"], c["max_terms"], m).queries
    k = c["k"]
    for opt in args.opts:
        ix._engine_cache.clear()
        try:
            s = Searcher(ix, weighting=BM25F, **parse_opts(opt))
        except Exception as e:                       # an option set the device cannot hold: report and go on
            print(json.dumps({"opt": opt, "error": str(e)}), flush=True)
            continue
        eng = s.engine
        for m, queries in qsets.items():
            batch = s.pack(queries)
            scores, docids, counts, totals = eng.search_batch(batch, k)
            step = max(1, len(queries) // max(1, args.check))
            bad = None
            for i in range(0, len(queries), step):
                n = int(counts[i])
                try:
                    assert_query_parity(o, queries[i], list(zip(scores[i, :n].tolist(), docids[i, :n].tolist())),
                                        int(totals[i]), k, ctx="%s query %d" % (opt, i))
                except AssertionError as e:
                    bad = str(e)
                    break
            plan = eng.prepare(batch, k, arena=args.arena)
            for _ in range(2):
                plan.execute()
            eng.synchronize()
            eng.reset_stats()
            for _ in range(args.steps):
                plan.execute()
            eng.synchronize()
            st = eng.stats()
            n = max(1, st["n_executes"])
            ms = st["ms_total"] / n
            gbs = 9.0 * st["postings_touched"] / (st["ms_score"] / n * 1e-3) / 1e9
            stream_gbs = 9.0 * st["postings_stream"] / max(1e-9, st["ms_stream"] / n * 1e-3) / 1e9
            print(json.dumps({"opt": opt, "mode": m, "parity": bad or "ok", "qps": len(queries) / (ms * 1e-3), "ms_total": ms,
                              "ms_bounds": st["ms_bounds"] / n, "ms_score": st["ms_score"] / n,
                              "ms_merge": st["ms_merge"] / n, "ms_stream": st["ms_stream"] / n,
                              "post_stream": st["postings_stream"], "post_lookup": st["postings_lookup"], "post_team": st["postings_team"],
                              "launches": st["n_launches"],
                              "stream_GBs": stream_gbs, "stream_frac_6547": stream_gbs / 6547.2,
                              "step_GBs": gbs, "items": st["n_items"], "ctas_per_sm": st["ctas_per_sm"],
                              "postings": st["postings_touched"]}), flush=True)
            eng.reset_stats()      # a BM25F_PROFILE build prints its phase timers here
            plan.close()
        eng.close()


if __name__ == "__main__":
    main()
