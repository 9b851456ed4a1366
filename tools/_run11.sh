timeout 300 python tools/r2_check.py check > gpurun_out/r2_check11.log 2>&1; echo "check rc=$?" >> gpurun_out/r2_check11.log
grep -E "BAD|CHECK|rc=|WATCHDOG|Error" gpurun_out/r2_check11.log
( for nc in 3 4 5; do PARADIS_SL_ROWS_NC=$nc timeout 100 python tools/r2_check.py time bilinear; done
timeout 100 python tools/r2_check.py time bicubic
PARADIS_SL_ROWS_NC=5 PARADIS_SL_LIB=build/variants/lib_w20t1.so timeout 100 python tools/r2_check.py time bilinear
for nc in 4 6; do PARADIS_SL_ROWS_NC=$nc PARADIS_SL_LIB=build/variants/lib_w24t2.so timeout 100 python tools/r2_check.py time bilinear; done
for nc in 3 4; do PARADIS_SL_ROWS_NC=$nc PARADIS_SL_LIB=build/variants/lib_w20t3.so timeout 100 python tools/r2_check.py time bilinear; done
for nc in 4 5; do PARADIS_SL_ROWS_NC=$nc PARADIS_SL_LIB=build/variants/lib_w24t3.so timeout 100 python tools/r2_check.py time bilinear; done
) > gpurun_out/r2_time11.log 2>&1
grep -E "TIME|WATCHDOG|Error" gpurun_out/r2_time11.log
