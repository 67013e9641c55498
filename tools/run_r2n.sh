set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -m gpu -q -x > gpurun_out/r2n_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_tests.log
timeout 600 python tools/knn_ab.py > gpurun_out/r2n_knn_ab.txt 2> gpurun_out/r2n_knn_ab.err
timeout 900 python tools/knn_full_parity.py > gpurun_out/r2n_knn_full_parity.txt 2> gpurun_out/r2n_knn_full_parity.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2n_knn_launches.csv python tools/ncu_knn_target.py 70000 128 11 > gpurun_out/r2n_knn_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_fused -s 4 -c 1 -o gpurun_out/r2n_knn_emit -f python tools/ncu_knn_target.py 70000 128 11 > gpurun_out/r2n_ncu_knn.log 2>&1
tail -8 gpurun_out/r2n_tests.log | cut -c1-200; cat gpurun_out/r2n_knn_ab.txt | cut -c1-220; cat gpurun_out/r2n_knn_full_parity.txt; tail -3 gpurun_out/r2n_knn_full_parity.err; tail -2 gpurun_out/r2n_ncu_knn.log
