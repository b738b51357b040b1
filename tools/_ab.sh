set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "contract or fused or golden or chunk or determin or ragged or full" 2>&1 | tail -5
B="python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-e2e --no-decode"
for w in c2 q1; do
  $B --workload $w > gpurun_out/ab_${w}_ew16.json 2>gpurun_out/ab_${w}_ew16.err
  QB_BLK32=1 $B --workload $w > gpurun_out/ab_${w}_ew16blk.json 2>/dev/null
  QINCO_B200_LIB=$PWD/qinco_b200/variants/ew8.so $B --workload $w > gpurun_out/ab_${w}_ew8.json 2>/dev/null
done
