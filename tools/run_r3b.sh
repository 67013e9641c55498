set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r3b_bench_8gpu.json 2> gpurun_out/r3b_bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3b_bench_8gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','n_gpus','ms_per_step')}, d['roofline']['frac'], d['e2e']['value'])
print(json.dumps(d.get('cfg5_rowpart'))[:900])
PY
tail -3 gpurun_out/r3b_bench_8gpu.err
