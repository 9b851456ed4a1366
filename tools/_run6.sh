timeout 300 python tools/r2_check.py check > gpurun_out/r2_check6.log 2>&1; echo "check rc=$?" >> gpurun_out/r2_check6.log
grep -E "BAD|CHECK|rc=|WATCHDOG|Error" gpurun_out/r2_check6.log
( timeout 100 python tools/r2_check.py time bilinear bicubic
for nc in 3 4 5 8; do PARADIS_SL_ROWS_NC=$nc timeout 100 python tools/r2_check.py time bilinear; done
for w in w16 w20; do for nc in 3 4 5; do PARADIS_SL_ROWS_NC=$nc PARADIS_SL_LIB=build/variants/lib_$w.so timeout 100 python tools/r2_check.py time bilinear; done; done ) > gpurun_out/r2_time6.log 2>&1
grep -E "TIME|WATCHDOG|Error" gpurun_out/r2_time6.log
