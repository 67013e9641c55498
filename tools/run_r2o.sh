set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -m gpu -q > gpurun_out/r2o_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2o_tests.log
timeout 900 python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
GLB_TIMING=1 timeout 300 python tools/first_fit.py > gpurun_out/r2o_first_fit.txt 2>&1
tail -25 gpurun_out/r2o_tests.log | cut -c1-220; cat gpurun_out/r2o_bench.json | cut -c1-6000; tail -3 gpurun_out/r2o_bench.err; grep "rep 1" -B14 gpurun_out/r2o_first_fit.txt | tail -16
