#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_poisson_gpu.py -x -q > $OUT/v3_poisson_tests.log 2>&1; echo "exit $?" >> $OUT/v3_poisson_tests.log
timeout 600 python tools/df_ab.py > $OUT/v3_df_ab.txt 2> $OUT/v3_df_ab.err; echo "exit $?" >> $OUT/v3_df_ab.err
timeout 400 python -m pytest tests/test_spectral_gpu.py -x -q > $OUT/v3_spectral_tests.log 2>&1; echo "exit $?" >> $OUT/v3_spectral_tests.log
timeout 300 python tools/spectral_probe.py > $OUT/v3_spectral_probe.json 2> $OUT/v3_spectral_probe.err; echo "exit $?" >> $OUT/v3_spectral_probe.err
for f in v3_poisson_tests.log v3_spectral_tests.log; do tail -n 4 $OUT/$f; done; cat $OUT/v3_df_ab.txt; tail -n 3 $OUT/v3_df_ab.err; cat $OUT/v3_spectral_probe.json
