PARADIS_SL_ROWS_NC=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sl_bwd_rows -c 1 -f -o gpurun_out/prof_rows1 python tools/prof_step.py 16 bilinear fast 1 6.0 > gpurun_out/ncu_rows1.log 2>&1
tail -3 gpurun_out/ncu_rows1.log
