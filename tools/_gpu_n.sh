set -x
N=$1
cd /root/repo; mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_plan_n$N.json 2> gpurun_out/r2_bench_plan_n$N.err; echo "bench rc=$?"
grep '^{' gpurun_out/r2_bench_plan_n$N.json | tail -c 4000; tail -5 gpurun_out/r2_bench_plan_n$N.err
