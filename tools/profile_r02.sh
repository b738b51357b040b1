#!/bin/bash
# Round-2 profiling pass (run under gpurun on ONE GPU; numbers printed under ncu are never bench values).
#   1. launch lists (cold-cache, serialised: compare SHARES) of a 65 536-vector C2 encode+decode and a 16 384-vector C3 encode
#   2. ncu --set full captures of the dominant kernels: fused score (resident), decode loop, tensor-core prep, fused score B, IVF
set -x
OUT=gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -s 46 -c 60 --csv --log-file $OUT/r02_launches_c2_n65536.csv $B --n 65536 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 45 -c 40 --csv --log-file $OUT/r02_launches_c3_n16384.csv $B --workload c3 --n 16384 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:qb_mlp_kernel -s 30 -c 2 -o $OUT/r02_prof_c2_score -f $B --n 65536 --no-decode > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:qb_prep_tc -s 8 -c 1 -o $OUT/r02_prof_c2a16_prep -f $B --workload c2a16 --n 65536 --no-decode > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:qb_mlp_kernel -s 2 -c 1 -o $OUT/r02_prof_c2_decode -f python tools/decode_probe.py c2 1000000 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:qb_mlp_kernel -s 8 -c 1 -o $OUT/r02_prof_c3_score -f $B --workload c3 --n 16384 --no-decode > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:qb_ivf_tc -s 2 -c 1 -o $OUT/r02_prof_ivf1m -f $B --workload ivf1m --n 18944 --no-decode > /dev/null 2>&1
ls -la $OUT/*.ncu-rep $OUT/r02_launches*
