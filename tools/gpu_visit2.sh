#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_plaplace_gpu.py -x -q > $OUT/v2_plaplace_tests.log 2>&1; echo "exit $?" >> $OUT/v2_plaplace_tests.log
timeout 200 python tools/plaplace_probe.py > $OUT/v2_plaplace_probe.json 2> $OUT/v2_plaplace_probe.err; echo "exit $?" >> $OUT/v2_plaplace_probe.err
timeout 400 python -m pytest tests/test_spectral_gpu.py -x -q > $OUT/v2_spectral_tests.log 2>&1; echo "exit $?" >> $OUT/v2_spectral_tests.log
timeout 300 python tools/spectral_probe.py > $OUT/v2_spectral_probe.json 2> $OUT/v2_spectral_probe.err; echo "exit $?" >> $OUT/v2_spectral_probe.err
for f in v2_plaplace_tests.log v2_spectral_tests.log; do tail -n 4 $OUT/$f; done; cat $OUT/v2_plaplace_probe.json $OUT/v2_spectral_probe.json; tail -n 5 $OUT/v2_spectral_probe.err
