set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_examples.py tests/test_mbo_gpu.py -m gpu -q > gpurun_out/r2y_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2y_tests.log
timeout 900 python bench.py --no-cfg5 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
tail -8 gpurun_out/r2y_tests.log | cut -c1-200; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['e2e']['value'])
o=d['other_rows']
print(json.dumps(o.get('batched_label_sets'), indent=1)); print(o.get('cfg3_laplace_60k_x_512_k20')); print(o.get('laplace_cg_fit')); print(o.get('error'))
PY
