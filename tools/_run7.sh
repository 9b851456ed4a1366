PARADIS_SL_ROWS_NC=5 PARADIS_SL_LIB=build/variants/lib_w20.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:sl_bwd_rows -c 1 -f -o gpurun_out/prof_rows2 python tools/prof_step.py 64 bilinear fast 1 6.0 > gpurun_out/ncu_rows2.log 2>&1
tail -2 gpurun_out/ncu_rows2.log
