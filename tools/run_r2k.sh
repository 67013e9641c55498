set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mbo_gpu.py tests/test_knn_gpu.py -m gpu -x -q > gpurun_out/r2k_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_knn_launches.csv python tools/ncu_knn_target.py 70000 128 11 > gpurun_out/r2k_knn_launches.log 2>&1
GLB_TIMING=1 timeout 300 python tools/first_fit.py > gpurun_out/r2k_first_fit.txt 2>&1
tail -12 gpurun_out/r2k_tests.log | cut -c1-200; grep -v "^==" gpurun_out/r2k_knn_launches.csv | awk -F'","' '{print $5, $NF}' | tail -30; grep "rep 2" -B14 gpurun_out/r2k_first_fit.txt
