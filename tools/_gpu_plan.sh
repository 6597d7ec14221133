cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_gputests.log
timeout 300 python bench.py --docs 125000 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_shard_emul.json 2> gpurun_out/r2_shard_emul.err; echo "bench rc=$?"
timeout 300 python bench.py --docs 250000 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_shard_emul4.json 2> gpurun_out/r2_shard_emul4.err; echo "bench rc=$?"
