#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 240 python -m pytest tests/test_poisson_gpu.py -x -q > $OUT/v5_poisson_tests.log 2>&1; echo "exit $?" >> $OUT/v5_poisson_tests.log
timeout 240 python tools/df_ab.py short > $OUT/v5_df_ab.txt 2> $OUT/v5_df_ab.err; echo "exit $?" >> $OUT/v5_df_ab.err
GLB_KIND=dataflow GLB_POISSON_NOPOLL=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:poisson_dataflow -s 8 -c 1 \
    -o $OUT/v5_dataflow2_nopoll python tools/ncu_target.py 100 2 > $OUT/v5_ncu_full.log 2>&1; echo "exit $?" >> $OUT/v5_ncu_full.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/v5_launches.csv \
    python bench.py --steps 2 --warmup 3 --iters 200 --cpu-iters 2 --no-extras > $OUT/v5_ncu_bench.log 2>&1; echo "exit $?" >> $OUT/v5_ncu_bench.log
timeout 400 python bench.py --steps 5 --warmup 3 > $OUT/v5_bench.json 2> $OUT/v5_bench.err; echo "exit $?" >> $OUT/v5_bench.err
tail -n 3 $OUT/v5_poisson_tests.log; cat $OUT/v5_df_ab.txt; tail -n 3 $OUT/v5_ncu_full.log $OUT/v5_ncu_bench.log; cut -c1-700 $OUT/v5_bench.json; ls -la $OUT | grep v5
