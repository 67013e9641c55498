set -x
mkdir -p gpurun_out
timeout 900 python tools/df_ab.py sleep > gpurun_out/r2b_df_sleep.txt 2> gpurun_out/r2b_df_sleep.err
cat gpurun_out/r2b_df_sleep.txt | tail -70
