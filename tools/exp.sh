for e in 0 1 2 3 4 8 12 16 32 63; do QB_EXP=$e QB_MLP_TRACE=gpurun_out/trace_exp$e.txt:12 python bench.py --n 65536 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1; done
