OUT=gpurun_out; mkdir -p $OUT
bash tools/gpu_full.sh s4final > $OUT/s4final_summary.txt 2>&1
timeout 300 python tools/cg_probe.py time > $OUT/s4final_cg_probe.txt 2>&1
timeout 300 python tools/default_fit_probe.py > $OUT/s4final_default_fit.txt 2>&1
timeout 300 python tools/first_fit.py > $OUT/s4final_first_fit.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/s4final_cg_launches.csv python tools/cg_probe.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cg_spmm_dot --launch-skip 50 --launch-count 1 -o $OUT/s4final_cg_spmm_70k -f python tools/cg_probe.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cg_spmm_dot --launch-skip 250 --launch-count 1 -o $OUT/s4final_cg_spmm_cfg3 -f python tools/cg_probe.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mixing_persistent --launch-skip 2 --launch-count 1 -o $OUT/s4final_mixing -f python tools/default_fit_probe.py > /dev/null 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/s4final_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cfg5 > /dev/null 2>&1
tail -12 $OUT/s4final_summary.txt; tail -4 $OUT/s4final_cg_probe.txt
