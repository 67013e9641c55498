#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/v9_knn_launches.csv \
    python tools/knn_ncu_target.py > $OUT/v9_knn_launches.log 2>&1; echo "exit $?" >> $OUT/v9_knn_launches.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:knn_dist_tc -s 3 -c 1 -o $OUT/v9_knn_dist_tc \
    python tools/knn_ncu_target.py > $OUT/v9_knn_ncu_full.log 2>&1; echo "exit $?" >> $OUT/v9_knn_ncu_full.log
timeout 300 ncu --set full --clock-control none -k regex:knn_select -s 3 -c 1 -o $OUT/v9_knn_select \
    python tools/knn_ncu_target.py > $OUT/v9_knn_ncu_select.log 2>&1; echo "exit $?" >> $OUT/v9_knn_ncu_select.log
tail -n 3 $OUT/v9_knn_launches.log $OUT/v9_knn_ncu_full.log; ls -la $OUT | grep v9
