#!/bin/bash
# Round-2 final pass on one B200: GPU tests, smoke, ncu captures, every bench line kept under profiles/.
set -x
OUT=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $OUT/r02_final_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/r02_final_smoke.txt 2>&1
timeout 400 bash tools/profile_r02.sh > $OUT/r02_final_profile.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/r02_bench_c2.json 2> $OUT/r02_bench_c2.err
for w in q1 c3 c4 c5 c2a16 livf ivf1m; do
  timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-extras > $OUT/r02_bench_$w.json 2> $OUT/r02_bench_$w.err
done
timeout 200 python bench.py --workload c5pw --steps 800 --warmup 3 > $OUT/r02_bench_c5pw.json 2> $OUT/r02_bench_c5pw.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/r02_bench_c2_reference.json 2> /dev/null
ls -la $OUT | tail -30
