set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -m gpu -q > gpurun_out/r2p_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2p_tests.log
timeout 900 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:poisson_dataflow -s 2 -c 1 -o gpurun_out/r2p_dataflow_T1000 -f python tools/ncu_target.py 1000 3 1 > gpurun_out/r2p_ncu_df.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cfg5 > gpurun_out/r2p_bench_under_ncu.log 2>&1
tail -12 gpurun_out/r2p_tests.log | cut -c1-220; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['e2e']['value'], d['e2e_default'])
o=d['other_rows']
for k in ('laplace_cg_fit','cfg3_laplace_60k_x_512_k20','error'):
    print(k, o.get(k))
PY
tail -3 gpurun_out/r2p_ncu_df.log; tail -3 gpurun_out/r2p_bench_under_ncu.log
