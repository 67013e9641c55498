#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -m pytest tests/test_knn_gpu.py -x -q > $OUT/v8_knn_tests.log 2>&1; echo "exit $?" >> $OUT/v8_knn_tests.log
GLB_KNN_TC=0 timeout 200 python tools/knn_ab.py > $OUT/v8_knn_fp32.txt 2>&1; echo "exit $?" >> $OUT/v8_knn_fp32.txt
GLB_KNN_TC=1 timeout 200 python tools/knn_ab.py > $OUT/v8_knn_tc.txt 2>&1; echo "exit $?" >> $OUT/v8_knn_tc.txt
tail -n 15 $OUT/v8_knn_tests.log; cat $OUT/v8_knn_fp32.txt $OUT/v8_knn_tc.txt
