set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -m gpu -q > gpurun_out/r2x_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2x_tests.log
timeout 600 python tools/df_ab.py threads > gpurun_out/r2x_df_threads.txt 2> gpurun_out/r2x_df_threads.err
tail -6 gpurun_out/r2x_tests.log | cut -c1-200; cat gpurun_out/r2x_df_threads.txt | cut -c1-200
