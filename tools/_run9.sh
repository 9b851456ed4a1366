timeout 300 python tools/r2_check.py check > gpurun_out/r2_check9.log 2>&1; echo "check rc=$?" >> gpurun_out/r2_check9.log
grep -E "BAD|CHECK|rc=|WATCHDOG|Error" gpurun_out/r2_check9.log
( for nc in 5 6; do PARADIS_SL_ROWS_NC=$nc timeout 100 python tools/r2_check.py time bilinear; done
for nc in 6 8; do PARADIS_SL_ROWS_NC=$nc PARADIS_SL_LIB=build/variants/lib_w24.so timeout 100 python tools/r2_check.py time bilinear; done
for nc in 7 9; do PARADIS_SL_ROWS_NC=$nc PARADIS_SL_LIB=build/variants/lib_w28.so timeout 100 python tools/r2_check.py time bilinear; done
for nc in 8 10 12; do PARADIS_SL_ROWS_NC=$nc PARADIS_SL_LIB=build/variants/lib_w32.so timeout 100 python tools/r2_check.py time bilinear; done
) > gpurun_out/r2_time9.log 2>&1
grep -E "TIME|WATCHDOG|Error" gpurun_out/r2_time9.log
