set -x
B="python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-e2e --no-decode"
QB_NO_FUSE=1 $B --workload c2 > gpurun_out/ab_c2_nofuse.json 2>/dev/null
$B --workload c2 --plan-opts pair=2 > gpurun_out/ab_c2_pair.json 2>gpurun_out/ab_c2_pair.err
$B --workload q1 --plan-opts pair=2 > gpurun_out/ab_q1_pair.json 2>/dev/null
$B --workload c3 --plan-opts pair=2 > gpurun_out/ab_c3_pair.json 2>/dev/null
