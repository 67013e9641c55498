set -x
mkdir -p gpurun_out
timeout 300 python tools/bench_cfg5.py --reps 3 --exchange put > gpurun_out/r2w_cfg5_base.json 2> gpurun_out/r2w_cfg5_base.err
GLB200_LIB=graphlearning_b200/lib/libglb200_exp.so timeout 300 python tools/bench_cfg5.py --reps 3 --exchange put > gpurun_out/r2w_cfg5_prefetch.json 2> gpurun_out/r2w_cfg5_prefetch.err
python - <<'PY'
import json
for f in ('base','prefetch'):
    d=json.loads(open('gpurun_out/r2w_cfg5_%s.json'%f).read().strip().splitlines()[-1])
    r=d['runs'][0]; print(f, r['ms_per_iteration'], r['iterations_per_s'], r['frac_of_hbm_peak_x_gpus'])
PY
