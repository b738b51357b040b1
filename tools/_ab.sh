timeout 300 python bench.py --workload c3 --vectors 1000000 --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-e2e --no-decode > gpurun_out/r02_bench_c3_1m.json 2> gpurun_out/r02_bench_c3_1m.err
timeout 400 python bench.py --workload c4 --vectors 1250000 --steps 1 --warmup 3 --no-extras --no-cpu-baseline --no-e2e --no-decode > gpurun_out/r02_bench_c4_1250k.json 2> gpurun_out/r02_bench_c4_1250k.err
tail -1 gpurun_out/r02_bench_c3_1m.json | cut -c1-250; tail -1 gpurun_out/r02_bench_c4_1250k.json | cut -c1-250; tail -2 gpurun_out/r02_bench_c4_1250k.err
