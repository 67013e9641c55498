#!/bin/bash
# GPU-box visit: new plaplace tests first (under a short timeout: a dataflow bug could spin), then the whole -m gpu suite,
# the plaplace probe, smoke, bench.
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_plaplace_gpu.py -x -q > $OUT/v1_plaplace_tests.log 2>&1; echo "exit $?" >> $OUT/v1_plaplace_tests.log
timeout 200 python tools/plaplace_probe.py > $OUT/v1_plaplace_probe.json 2> $OUT/v1_plaplace_probe.err; echo "exit $?" >> $OUT/v1_plaplace_probe.err
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_plaplace_gpu.py > $OUT/v1_tests.log 2>&1; echo "exit $?" >> $OUT/v1_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/v1_smoke.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/v1_bench.json 2> $OUT/v1_bench.err
tail -3 $OUT/v1_plaplace_tests.log $OUT/v1_tests.log $OUT/v1_smoke.log; cat $OUT/v1_plaplace_probe.json
