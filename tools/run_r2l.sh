set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -m gpu -x -q > gpurun_out/r2l_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2l_tests.log
timeout 600 python tools/knn_ab.py > gpurun_out/r2l_knn_ab.txt 2> gpurun_out/r2l_knn_ab.err
GLB_TIMING=1 timeout 300 python tools/first_fit.py > gpurun_out/r2l_first_fit.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2l_knn_launches.csv python tools/ncu_knn_target.py 70000 128 11 > gpurun_out/r2l_knn_launches.log 2>&1
tail -8 gpurun_out/r2l_tests.log | cut -c1-200; cat gpurun_out/r2l_knn_ab.txt; grep "rep 2" -B12 gpurun_out/r2l_first_fit.txt
