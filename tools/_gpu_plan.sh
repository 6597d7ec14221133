cd /root/repo; mkdir -p gpurun_out
show='
import sys, json
for l in sys.stdin:
    d=json.loads(l); print({k: d[k] for k in ("opt","mode","parity","ms_total","ms_score","ms_stream","ms_merge","items","post_stream","post_lookup")})'
timeout 500 python tools/tune.py --config 4 --queries 50000 --steps 3 --check 8 --opts default isect_ratio=2 isect_ratio=4 isect_or_limit=10000 isect_or_limit=160000 serial_streams=1 2> gpurun_out/tune.err | python -c "$show"
tail -2 gpurun_out/tune.err
