set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2m_gpu.txt
timeout 900 python -m pytest tests/ -m gpu -q > gpurun_out/r2m_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_tests.log
timeout 900 python tools/df_ab.py octet > gpurun_out/r2m_df_octet.txt 2> gpurun_out/r2m_df_octet.err
timeout 600 python tools/df_ab.py octet hub > gpurun_out/r2m_df_octet_hub.txt 2> gpurun_out/r2m_df_octet_hub.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:poisson_dataflow -s 1 -c 1 -o gpurun_out/r2m_dataflow_octet -f python tools/ncu_target.py 100 3 1 > gpurun_out/r2m_ncu_df.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_fused -c 4 -o gpurun_out/r2m_knn_fused -f python tools/ncu_knn_target.py 70000 128 11 > gpurun_out/r2m_ncu_knn.log 2>&1
tail -15 gpurun_out/r2m_tests.log | cut -c1-200; cat gpurun_out/r2m_df_octet.txt; cat gpurun_out/r2m_df_octet_hub.txt; tail -3 gpurun_out/r2m_ncu_df.log; tail -3 gpurun_out/r2m_ncu_knn.log
