set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_tests.log
timeout 600 python tools/bench_cfg5.py --reps 3 --exchange put > gpurun_out/r2g_cfg5_1gpu.json 2> gpurun_out/r2g_cfg5_1gpu.err
timeout 900 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
tail -8 gpurun_out/r2g_tests.log; cat gpurun_out/r2g_cfg5_1gpu.json; tail -n 3 gpurun_out/r2g_cfg5_1gpu.err; cat gpurun_out/r2g_bench.json; tail -n 5 gpurun_out/r2g_bench.err
