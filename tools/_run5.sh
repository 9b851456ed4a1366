timeout 300 python tools/r2_check.py check > gpurun_out/r2_check5.log 2>&1; echo "check rc=$?" >> gpurun_out/r2_check5.log
grep -E "BAD|CHECK|rc=|WATCHDOG|Error" gpurun_out/r2_check5.log
( timeout 100 python tools/r2_check.py time bilinear bicubic
for nc in 4 5 8; do PARADIS_SL_ROWS_NC=$nc timeout 100 python tools/r2_check.py time bilinear; done
for w in w20 w28; do for nc in 5 7; do PARADIS_SL_ROWS_NC=$nc PARADIS_SL_LIB=build/variants/lib_$w.so timeout 100 python tools/r2_check.py time bilinear; done; done ) > gpurun_out/r2_time5.log 2>&1
grep -E "TIME|WATCHDOG|Error" gpurun_out/r2_time5.log
