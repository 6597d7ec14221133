cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_gputests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
