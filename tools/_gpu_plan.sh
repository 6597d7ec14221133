cd /root/repo; mkdir -p gpurun_out
show='
import sys, json
for l in sys.stdin:
    d=json.loads(l); print({k: d[k] for k in ("opt","mode","parity","ms_total","ms_score","ms_stream","items","post_stream","post_lookup")})'
export BM25F_LIB=/root/repo/document_search_engine_b200/csrc/libbm25f_old.so
timeout 300 python tools/tune.py --config 2 --steps 10 --modes cfg and --opts default 2> gpurun_out/tune.err | python -c "$show"
timeout 300 python tools/tune.py --config 2 --docs 125000 --steps 10 --modes cfg and --opts default 2> gpurun_out/tune.err | python -c "$show"
timeout 300 python tools/tune.py --config 3 --steps 10 --opts default 2> gpurun_out/tune.err | python -c "$show"
