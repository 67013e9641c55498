#!/bin/bash
# every step under a SHORT timeout: a polling kernel that never finishes must not eat the budget
OUT=gpurun_out; mkdir -p $OUT
timeout 240 python -m pytest tests/test_poisson_gpu.py -x -q > $OUT/v4_poisson_tests.log 2>&1; echo "exit $?" >> $OUT/v4_poisson_tests.log
timeout 300 python tools/df_ab.py > $OUT/v4_df_ab.txt 2> $OUT/v4_df_ab.err; echo "exit $?" >> $OUT/v4_df_ab.err
timeout 150 python -m pytest tests/test_plaplace_gpu.py -x -q > $OUT/v4_plaplace_tests.log 2>&1; echo "exit $?" >> $OUT/v4_plaplace_tests.log
timeout 150 python tools/plaplace_probe.py > $OUT/v4_plaplace_probe.json 2> $OUT/v4_plaplace_probe.err; echo "exit $?" >> $OUT/v4_plaplace_probe.err
timeout 200 python -m pytest tests/test_spectral_gpu.py -x -q > $OUT/v4_spectral_tests.log 2>&1; echo "exit $?" >> $OUT/v4_spectral_tests.log
timeout 150 python tools/spectral_probe.py > $OUT/v4_spectral_probe.json 2> $OUT/v4_spectral_probe.err; echo "exit $?" >> $OUT/v4_spectral_probe.err
for f in v4_poisson_tests.log v4_plaplace_tests.log v4_spectral_tests.log; do tail -n 4 $OUT/$f; done; cat $OUT/v4_df_ab.txt; tail -n 3 $OUT/v4_df_ab.err; cat $OUT/v4_plaplace_probe.json $OUT/v4_spectral_probe.json
