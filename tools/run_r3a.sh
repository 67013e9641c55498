set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -m gpu -q -x > gpurun_out/r3a_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r3a_tests.log
timeout 900 python bench.py > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3a_bench_ref.json 2> gpurun_out/r3a_bench_ref.err
tail -6 gpurun_out/r3a_tests.log | cut -c1-300; cut -c1-400 gpurun_out/r3a_bench.json; cut -c1-600 gpurun_out/r3a_bench_ref.json
