set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2h_topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29711 tools/bench_cfg5.py --reps 3 --exchange both > gpurun_out/r2h_cfg5_4gpu.json 2> gpurun_out/r2h_cfg5_4gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2h_bench_8gpu.json 2> gpurun_out/r2h_bench_8gpu.err
cat gpurun_out/r2h_cfg5_4gpu.json; tail -n 3 gpurun_out/r2h_cfg5_4gpu.err; cat gpurun_out/r2h_bench_8gpu.json; tail -n 5 gpurun_out/r2h_bench_8gpu.err
