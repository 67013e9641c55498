set -x
mkdir -p gpurun_out
timeout 600 python tools/wstats_probe.py > gpurun_out/r2s_wstats.log 2>&1
tail -4 gpurun_out/r2s_wstats.log; head -3 gpurun_out/wstats.txt
