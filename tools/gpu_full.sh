#!/bin/bash
# Full check on one GPU: the whole -m gpu suite, smoke, bench (both arms).  Usage: bash tools/gpu_full.sh <tag>
TAG=${1:-full}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; echo "exit $?" >> $OUT/${TAG}_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "exit $?" >> $OUT/${TAG}_smoke.log
timeout 500 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "exit $?" >> $OUT/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err; echo "exit $?" >> $OUT/${TAG}_bench.err
tail -n 4 $OUT/${TAG}_tests.log $OUT/${TAG}_smoke.log; tail -n 4 $OUT/${TAG}_bench.err; cut -c1-400 $OUT/${TAG}_bench.json; cut -c1-300 $OUT/${TAG}_bench_ref.json
