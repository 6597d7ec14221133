# Round-2 measurement script (run through gpurun): see profiles/r02_* for what it produced.
cd /root/repo; mkdir -p gpurun_out
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err; echo "ref rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_n1.csv python bench.py --steps 2 --warmup 1 --no-cpu --check 0 > gpurun_out/ncu1.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_score_stream|k_score_isect" -c 4 -f -o gpurun_out/r02_full python bench.py --steps 1 --warmup 1 --no-cpu --check 0 > gpurun_out/ncu2.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/r02_full.ncu-rep
timeout 300 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_bench_config3_n1.json 2> gpurun_out/r02_bench_config3_n1.err; echo "config3 rc=$?"
