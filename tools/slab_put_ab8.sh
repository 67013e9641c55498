#!/bin/bash
# the 8-GPU subset of tools/slab_put_ab.sh
N=8; OUT=gpurun_out/slab_put_ab_${N}gpu.txt; : > $OUT
export GLB200_LIB=$PWD/graphlearning_b200/lib/libglb200_exp.so
run() {
  echo "== $1" >> $OUT
  env $1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/bench_cfg5.py --exchange put 2>/dev/null | grep -o '"ms_per_iteration": [0-9.]*\|"iterations_per_s": [0-9.]*' | tr '\n' ' ' >> $OUT; echo >> $OUT
}
run "GLB_SLAB_EXP=2"
run "GLB_SLAB_EXP=0"
run "GLB_SLAB_EXP=1"
run "GLB_SLAB_EXP=0 GLB_SLAB_BND_FRAC=0.75"
run "GLB_SLAB_EXP=0 GLB_SLAB_BND_FRAC=1.0"
cat $OUT
