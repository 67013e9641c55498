set -x
mkdir -p gpurun_out
timeout 300 python tools/bench_cfg5.py --reps 3 --exchange put > gpurun_out/r2q_cfg5_minb3.json 2> gpurun_out/r2q_cfg5_minb3.err
GLB200_LIB=graphlearning_b200/lib/libglb200_exp.so timeout 300 python tools/bench_cfg5.py --reps 3 --exchange put > gpurun_out/r2q_cfg5_minb4.json 2> gpurun_out/r2q_cfg5_minb4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:slab_step -s 3 -c 1 -o gpurun_out/r2q_slab_step -f python tools/ncu_slab_target.py 2000000 6 > gpurun_out/r2q_ncu_slab.log 2>&1
cut -c1-700 gpurun_out/r2q_cfg5_minb3.json; cut -c1-700 gpurun_out/r2q_cfg5_minb4.json; tail -2 gpurun_out/r2q_ncu_slab.log
