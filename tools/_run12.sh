( for i in 1 2; do timeout 100 python tools/r2_check.py time bilinear; done
for v in g2r2 g3r2 g3r3 g4r2; do PARADIS_SL_LIB=build/variants/lib_$v.so timeout 100 python tools/r2_check.py time bilinear; done
for nc in 6 8 10; do PARADIS_SL_ROWS_NC=$nc timeout 100 python tools/r2_check.py time bicubic; done
PARADIS_SL_BWD=1 timeout 100 python tools/r2_check.py time bicubic
) > gpurun_out/r2_time12.log 2>&1
grep -E "TIME|WATCHDOG|Error" gpurun_out/r2_time12.log
