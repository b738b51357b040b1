OUT=gpurun_out
for w in c3 q1 c4 c5; do
  timeout 110 python bench.py --workload $w --steps 3 --warmup 3 --no-extras > $OUT/r02_bench_$w.json 2> $OUT/r02_bench_$w.err
done
for w in c3 q1 c4 c5; do tail -1 $OUT/r02_bench_$w.json | cut -c1-120; done
