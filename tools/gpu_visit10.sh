#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -m pytest tests/test_knn_gpu.py -x -q > $OUT/v10_knn_tests.log 2>&1; echo "exit $?" >> $OUT/v10_knn_tests.log
GLB_KNN_TC=1 timeout 200 python tools/knn_ab.py > $OUT/v10_knn_tc.txt 2>&1; echo "exit $?" >> $OUT/v10_knn_tc.txt
GLB_KNN_TC=1 GLB_KNN_QB=2048 timeout 200 python tools/knn_ab.py > $OUT/v10_knn_tc_qb2048.txt 2>&1; echo "exit $?" >> $OUT/v10_knn_tc_qb2048.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/v10_knn_launches.csv \
    python tools/knn_ncu_target.py > $OUT/v10_knn_launches.log 2>&1; echo "exit $?" >> $OUT/v10_knn_launches.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:knn_dist_tc -s 1 -c 1 -o $OUT/v10_knn_dist_tc \
    python tools/knn_ncu_target.py > $OUT/v10_knn_ncu_full.log 2>&1; echo "exit $?" >> $OUT/v10_knn_ncu_full.log
tail -n 5 $OUT/v10_knn_tests.log; cat $OUT/v10_knn_tc.txt $OUT/v10_knn_tc_qb2048.txt
