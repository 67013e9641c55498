set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests/test_poisson_gpu.py -m gpu -x -q > gpurun_out/r2a_poisson_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_poisson_tests.log
timeout 600 python tools/df_ab.py > gpurun_out/r2a_df_ab.txt 2> gpurun_out/r2a_df_ab.err
timeout 600 python tools/df_ab.py hub > gpurun_out/r2a_df_ab_hub.txt 2> gpurun_out/r2a_df_ab_hub.err
timeout 300 python tools/bench_cfg5.py --reps 2 > gpurun_out/r2a_cfg5_nat.json 2> gpurun_out/r2a_cfg5_nat.err
timeout 300 python tools/bench_cfg5.py --reps 2 --reorder 1 > gpurun_out/r2a_cfg5_rcm.json 2> gpurun_out/r2a_cfg5_rcm.err
tail -3 gpurun_out/r2a_poisson_tests.log; cat gpurun_out/r2a_df_ab.txt; cat gpurun_out/r2a_df_ab_hub.txt; cat gpurun_out/r2a_cfg5_*.json
