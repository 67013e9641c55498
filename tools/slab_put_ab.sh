#!/bin/bash
# A/B of the boundary phase of slab_step_kernel on N GPUs (experiment build): fence per tile (round-2 behaviour until session 4)
# vs one fence per CTA, puts skipped (cost probe, results wrong), share of the CTAs that start on the boundary tiles.
# usage: bash tools/slab_put_ab.sh N
N=${1:-2}; OUT=gpurun_out/slab_put_ab_${N}gpu.txt; : > $OUT
export GLB200_LIB=$PWD/graphlearning_b200/lib/libglb200_exp.so
run() {
  echo "== $1" >> $OUT
  env $1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/bench_cfg5.py --exchange put 2>/dev/null | grep -o '"ms_per_iteration": [0-9.]*\|"iterations_per_s": [0-9.]*' | tr '\n' ' ' >> $OUT; echo >> $OUT
}
run "GLB_SLAB_EXP=2"
run "GLB_SLAB_EXP=0"
run "GLB_SLAB_EXP=1"
run "GLB_SLAB_EXP=0 GLB_SLAB_BND_FRAC=0.3"
run "GLB_SLAB_EXP=0 GLB_SLAB_BND_FRAC=0.75"
run "GLB_SLAB_EXP=0 GLB_SLAB_BND_FRAC=1.0"
cat $OUT
