set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_poisson_gpu.py tests/test_cg_gpu.py tests/test_laplace_gpu.py -m gpu -q -x > gpurun_out/r2t_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2t_tests.log
timeout 600 python tools/df_ab.py > gpurun_out/r2t_df_ab.txt 2> gpurun_out/r2t_df_ab.err
timeout 600 python tools/df_ab.py hub > gpurun_out/r2t_df_ab_hub.txt 2> gpurun_out/r2t_df_ab_hub.err
timeout 600 python tools/wstats_probe.py > gpurun_out/r2t_wstats.log 2>&1
tail -5 gpurun_out/r2t_tests.log | cut -c1-200; head -4 gpurun_out/r2t_df_ab.txt | cut -c1-200; head -4 gpurun_out/r2t_df_ab_hub.txt | cut -c1-200; grep "glb\]" gpurun_out/r2t_wstats.log | tail -2
