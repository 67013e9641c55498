#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python tools/first_fit.py > $OUT/v12_first_fit.txt 2>&1; echo "exit $?" >> $OUT/v12_first_fit.txt
timeout 200 python -m pytest tests/test_cg_gpu.py -x -q > $OUT/v12_cg_tests.log 2>&1; echo "exit $?" >> $OUT/v12_cg_tests.log
cat $OUT/v12_first_fit.txt; tail -n 5 $OUT/v12_cg_tests.log
