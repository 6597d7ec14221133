set -x
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_gputests.log
