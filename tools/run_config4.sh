# Round-2 measurement script (run through gpurun): see profiles/r02_* for what it produced.
N=$1
cd /root/repo; mkdir -p gpurun_out
if [ $N = 1 ]; then
timeout 900 python bench.py --config 4 --queries 50000 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench_config4_n$N.json 2> gpurun_out/r02_bench_config4_n$N.err; echo "bench rc=$?"
else
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config 4 --queries 50000 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench_config4_n$N.json 2> gpurun_out/r02_bench_config4_n$N.err; echo "bench rc=$?"
fi
tail -3 gpurun_out/r02_bench_config4_n$N.err
grep '^{' gpurun_out/r02_bench_config4_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['serial_value'], d['kernel_ms'], d['roofline']['frac'], d['config']['engine'])"
