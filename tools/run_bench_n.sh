# Round-2 measurement script (run through gpurun): see profiles/r02_* for what it produced.
N=$1
cd /root/repo; mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"
grep '^{' gpurun_out/r02_bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['serial_value'], d['kernel_ms'])"
