cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_plan.py tests/test_gpu_edges.py -x -q -m gpu 2>&1 | tail -3
BM25F_TRACE=1 timeout 300 python bench.py --docs 125000 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_shard_emul.json 2> gpurun_out/r2_shard_emul.err; echo "bench rc=$?"
grep "device planner" gpurun_out/r2_shard_emul.err | sed -n '30,34p'
