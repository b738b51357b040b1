set -x
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --no-e2e --no-decode"
for w in c2 q1 c3; do
  $B --workload $w > gpurun_out/ab_${w}_new.json 2>gpurun_out/ab_${w}_new.err
  QB_NO_BLK32=1 $B --workload $w > gpurun_out/ab_${w}_noblk.json 2>/dev/null
  QINCO_B200_LIB=$PWD/qinco_b200/variants/lib_arrive_elect.so $B --workload $w > gpurun_out/ab_${w}_elect.json 2>/dev/null
done
grep -h -o '"value": [0-9.]*\|"frac": [0-9.]*' gpurun_out/ab_*.json | head -0
for f in gpurun_out/ab_*.json; do echo $f; python -c "
import json,sys
for l in open('$f'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['roofline']['frac'], d['roofline'].get('avg_launch_ms'))
"; done
