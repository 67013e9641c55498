#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list + one full capture of the top kernel.
# Usage (from the repo root on the box): bash tools/gpu_round.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --ref-iters 50 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
# every launch of a short bench run with its device time (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --iters 200 --cpu-iters 2 > $OUT/${TAG}_ncu_bench.log 2>&1
# the top kernel, full set, with source
timeout 900 ncu --set full --clock-control none --import-source on -k regex:poisson_persistent -s 3 -c 1 \
    -o $OUT/${TAG}_persistent python bench.py --steps 2 --warmup 3 --iters 200 --cpu-iters 2 > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT
