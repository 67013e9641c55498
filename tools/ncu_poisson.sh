#!/bin/bash
# full ncu capture of the persistent Poisson kernel on the bench graph (one GPU)
# usage: bash tools/ncu_poisson.sh <tag> [variant]
TAG=${1:-r1}
export GLB_POISSON_VARIANT=${2:-512,16}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:poisson_dataflow -s 3 -c 1 \
    -o gpurun_out/${TAG}_dataflow python bench.py --steps 2 --warmup 3 --iters 200 --cpu-iters 2 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
