set -x
mkdir -p gpurun_out
export EXP=graphlearning_b200/lib/libglb200_exp.so
timeout 600 python tools/df_ab.py free > gpurun_out/r2c_df_free.txt 2> gpurun_out/r2c_df_free.err
timeout 600 python -m pytest tests/test_distributed_gpu.py -m gpu -x -q > gpurun_out/r2c_dist_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_dist_tests.log
timeout 600 python tools/bench_cfg5.py --reps 2 --exchange both > gpurun_out/r2c_cfg5_1gpu.json 2> gpurun_out/r2c_cfg5_1gpu.err
timeout 300 python tools/bench_cfg5.py --reps 2 --exchange put --reorder 0 > gpurun_out/r2c_cfg5_1gpu_nat.json 2> gpurun_out/r2c_cfg5_1gpu_nat.err
GLB200_LIB=$EXP GLB_KIND=dataflow GLB_POISSON_GATE_EVERY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:poisson_dataflow -s 2 -c 1 -o gpurun_out/r2c_pipe_gate1 -f python tools/ncu_target.py 100 4 1 > gpurun_out/r2c_ncu_pipe.log 2>&1
GLB200_LIB=$EXP GLB_KIND=dataflow GLB_POISSON_GATE_EVERY=0 GLB_POISSON_FREE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:poisson_dataflow -s 2 -c 1 -o gpurun_out/r2c_pipe_free -f python tools/ncu_target.py 100 4 1 > gpurun_out/r2c_ncu_pipe_free.log 2>&1
cut -c1-140 gpurun_out/r2c_df_free.txt; tail -15 gpurun_out/r2c_dist_tests.log; cat gpurun_out/r2c_cfg5_1gpu.json gpurun_out/r2c_cfg5_1gpu_nat.json; tail -3 gpurun_out/r2c_cfg5_1gpu.err; tail -2 gpurun_out/r2c_ncu_pipe.log gpurun_out/r2c_ncu_pipe_free.log
