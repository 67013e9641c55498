set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_knn_gpu.py tests/test_distributed_gpu.py tests/test_examples.py -m gpu -x -q > gpurun_out/r2i_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_tests.log
timeout 900 python tools/knn_ab.py full > gpurun_out/r2i_knn_ab.txt 2> gpurun_out/r2i_knn_ab.err
timeout 300 python tools/bench_cfg5.py --reps 3 --exchange put > gpurun_out/r2i_cfg5_1gpu.json 2> gpurun_out/r2i_cfg5_1gpu.err
GLB_TIMING=1 timeout 300 python tools/first_fit.py > gpurun_out/r2i_first_fit.txt 2>&1
tail -12 gpurun_out/r2i_tests.log; cat gpurun_out/r2i_knn_ab.txt; tail -n 3 gpurun_out/r2i_knn_ab.err; cat gpurun_out/r2i_cfg5_1gpu.json; grep -v "^$" gpurun_out/r2i_first_fit.txt | tail -60
