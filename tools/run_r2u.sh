set -x
mkdir -p gpurun_out
timeout 600 python tools/df_ab.py > gpurun_out/r2u_df_ab.txt 2> gpurun_out/r2u_df_ab.err
timeout 600 python tools/wstats_probe.py > gpurun_out/r2u_wstats.log 2>&1
head -3 gpurun_out/r2u_df_ab.txt | cut -c1-200; grep "glb\]" gpurun_out/r2u_wstats.log | tail -2
