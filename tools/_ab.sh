timeout 840 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -x -q > gpurun_out/r02_memcheck_tests.txt 2>&1
echo "exit code $?" >> gpurun_out/r02_memcheck_tests.txt
tail -8 gpurun_out/r02_memcheck_tests.txt
