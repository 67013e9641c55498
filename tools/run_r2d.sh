set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_distributed_gpu.py -m gpu -x -q > gpurun_out/r2d_dist_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_dist_tests.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:slab_step -s 3 -c 1 -o gpurun_out/r2d_slab_step -f python tools/ncu_slab_target.py 2000000 6 > gpurun_out/r2d_ncu_slab.log 2>&1
tail -8 gpurun_out/r2d_dist_tests.log; tail -n 4 gpurun_out/r2d_ncu_slab.log
