#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 150 python -m pytest tests/test_plaplace_gpu.py -x -q > $OUT/v6_plaplace_tests.log 2>&1; echo "exit $?" >> $OUT/v6_plaplace_tests.log
timeout 150 python tools/plaplace_probe.py > $OUT/v6_plaplace_probe.json 2> $OUT/v6_plaplace_probe.err; echo "exit $?" >> $OUT/v6_plaplace_probe.err
GLB_KIND=dataflow GLB_POISSON_NOPOLL=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:poisson_dataflow -s 2 -c 1 \
    -o $OUT/v6_dataflow2_nopoll python tools/ncu_target.py 100 2 > $OUT/v6_ncu_full.log 2>&1; echo "exit $?" >> $OUT/v6_ncu_full.log
GLB_BENCH_KIND=dataflow timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/v6_launches.csv \
    python bench.py --steps 2 --warmup 3 --iters 200 --cpu-iters 2 --no-extras > $OUT/v6_ncu_bench.log 2>&1; echo "exit $?" >> $OUT/v6_ncu_bench.log
tail -n 3 $OUT/v6_plaplace_tests.log; cat $OUT/v6_plaplace_probe.json; tail -n 4 $OUT/v6_ncu_full.log; ls -la $OUT | grep v6
