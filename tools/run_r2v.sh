set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2v_gpus.txt
timeout 900 python -m pytest tests/test_distributed_gpu.py -m gpu -q > gpurun_out/r2v_dist_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2v_dist_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2v_bench_2gpu.json 2> gpurun_out/r2v_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2v_bench_2gpu_ref.json 2> gpurun_out/r2v_bench_2gpu_ref.err
tail -4 gpurun_out/r2v_dist_tests.log; python - <<'PY'
import json
for f in ('gpurun_out/r2v_bench_2gpu.json','gpurun_out/r2v_bench_2gpu_ref.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('impl','value','n_gpus','ms_per_step')}, (d.get('roofline') or {}).get('frac'), json.dumps(d.get('cfg5_rowpart'))[:700])
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/r2v_bench_2gpu.err
