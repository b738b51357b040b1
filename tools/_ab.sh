T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2"
timeout 200 $T --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-decode > gpurun_out/r02_bench_c2_n2.json 2> gpurun_out/r02_bench_c2_n2.err
timeout 300 $T --workload c4 --vectors 300000 --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-e2e --no-decode > gpurun_out/r02_bench_c4_n2.json 2> gpurun_out/r02_bench_c4_n2.err
tail -1 gpurun_out/r02_bench_c2_n2.json | cut -c1-300; tail -1 gpurun_out/r02_bench_c4_n2.json | cut -c1-300
