B="python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline --no-e2e --no-decode"
$B --workload c3 > gpurun_out/ab_c3_fixed.json 2>gpurun_out/ab_c3_fixed.err
QB_NO_FIXED_SHAPE=1 $B --workload c3 > gpurun_out/ab_c3_generic.json 2>/dev/null
$B --workload c5 > gpurun_out/ab_c5_fixed.json 2>gpurun_out/ab_c5_fixed.err
$B --workload c4 > gpurun_out/ab_c4_fixed.json 2>gpurun_out/ab_c4_fixed.err
