set -x
mkdir -p gpurun_out
GLB_POISSON_STATS=1 timeout 600 python tools/df_ab.py stats > gpurun_out/r2r_df_stats.txt 2> gpurun_out/r2r_df_stats.err
cat gpurun_out/r2r_df_stats.txt | cut -c1-200; grep "glb\]" gpurun_out/r2r_df_stats.err | tail -24
