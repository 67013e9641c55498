#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
GLB_LIP_MODE=8 timeout 120 python -m pytest tests/test_plaplace_gpu.py -x -q > $OUT/v17_tests_mode8.log 2>&1; echo "exit $?" >> $OUT/v17_tests_mode8.log
GLB_LIP_MODE=15 timeout 120 python -m pytest tests/test_plaplace_gpu.py -x -q > $OUT/v17_tests_mode15.log 2>&1; echo "exit $?" >> $OUT/v17_tests_mode15.log
timeout 150 python tools/lip_modes.py > $OUT/v17_lip_modes.txt 2> $OUT/v17_lip_modes.err; echo "exit $?" >> $OUT/v17_lip_modes.err
tail -n 6 $OUT/v17_tests_mode8.log $OUT/v17_tests_mode15.log; cat $OUT/v17_lip_modes.txt; tail -n 3 $OUT/v17_lip_modes.err
