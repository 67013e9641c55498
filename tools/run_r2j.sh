set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_distributed_gpu.py tests/test_poisson_gpu.py -m gpu -x -q > gpurun_out/r2j_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_tests.log
timeout 300 python tools/bench_cfg5.py --reps 3 --exchange put > gpurun_out/r2j_cfg5_1gpu.json 2> gpurun_out/r2j_cfg5_1gpu.err
GLB_TIMING=1 timeout 300 python tools/first_fit.py > gpurun_out/r2j_first_fit.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_fused -c 3 -o gpurun_out/r2j_knn_fused -f python tools/ncu_knn_target.py 37888 128 11 > gpurun_out/r2j_ncu_knn.log 2>&1
tail -6 gpurun_out/r2j_tests.log; cat gpurun_out/r2j_cfg5_1gpu.json | cut -c1-600; grep "rep 1" -B16 gpurun_out/r2j_first_fit.txt | tail -18; tail -n 4 gpurun_out/r2j_ncu_knn.log
