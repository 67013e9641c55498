#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 150 python -m pytest tests/test_plaplace_gpu.py -x -q > $OUT/v7_plaplace_tests.log 2>&1; echo "exit $?" >> $OUT/v7_plaplace_tests.log
timeout 200 python tools/lip_modes.py > $OUT/v7_lip_modes.txt 2> $OUT/v7_lip_modes.err; echo "exit $?" >> $OUT/v7_lip_modes.err
tail -n 3 $OUT/v7_plaplace_tests.log; cat $OUT/v7_lip_modes.txt; tail -n 3 $OUT/v7_lip_modes.err
