( for v in w24s1 w24s2; do for nc in 6 8; do PARADIS_SL_ROWS_NC=$nc PARADIS_SL_LIB=build/variants/lib_$v.so timeout 100 python tools/r2_check.py time bilinear; done; done
for v in w32s1 w32s2; do for nc in 8 10; do PARADIS_SL_ROWS_NC=$nc PARADIS_SL_LIB=build/variants/lib_$v.so timeout 100 python tools/r2_check.py time bilinear; done; done
) > gpurun_out/r2_time10.log 2>&1
grep -E "TIME|WATCHDOG|Error" gpurun_out/r2_time10.log
