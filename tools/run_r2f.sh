set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2f_topo.txt 2>&1
timeout 900 python -m pytest tests/test_distributed_gpu.py -m gpu -x -q > gpurun_out/r2f_dist_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_dist_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 tools/bench_cfg5.py --reps 3 --exchange both > gpurun_out/r2f_cfg5_2gpu.json 2> gpurun_out/r2f_cfg5_2gpu.err
tail -12 gpurun_out/r2f_dist_tests.log; cat gpurun_out/r2f_cfg5_2gpu.json; tail -n 5 gpurun_out/r2f_cfg5_2gpu.err
