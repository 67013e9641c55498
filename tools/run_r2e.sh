set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_poisson_gpu.py tests/test_distributed_gpu.py -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_tests.log
timeout 600 python tools/df_ab.py > gpurun_out/r2e_df_ab.txt 2> gpurun_out/r2e_df_ab.err
timeout 600 python tools/df_ab.py hub > gpurun_out/r2e_df_ab_hub.txt 2> gpurun_out/r2e_df_ab_hub.err
timeout 600 python tools/bench_cfg5.py --reps 2 --exchange put > gpurun_out/r2e_cfg5_1gpu.json 2> gpurun_out/r2e_cfg5_1gpu.err
tail -8 gpurun_out/r2e_tests.log; cut -c1-150 gpurun_out/r2e_df_ab.txt; cut -c1-150 gpurun_out/r2e_df_ab_hub.txt; cat gpurun_out/r2e_cfg5_1gpu.json; tail -n 3 gpurun_out/r2e_cfg5_1gpu.err
