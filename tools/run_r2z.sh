set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_examples.py tests/test_mbo_gpu.py -m gpu -q > gpurun_out/r2z_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2z_tests.log
tail -8 gpurun_out/r2z_tests.log | cut -c1-300
