timeout 300 python tools/r2_check.py check > gpurun_out/r2_check3.log 2>&1; echo "check rc=$?" >> gpurun_out/r2_check3.log
tail -25 gpurun_out/r2_check3.log
( timeout 100 python tools/r2_check.py time bilinear bicubic
for nc in 4 8 10; do PARADIS_SL_ROWS_NC=$nc timeout 100 python tools/r2_check.py time bilinear; done ) > gpurun_out/r2_time3.log 2>&1
grep -E "TIME|WATCHDOG|Error" gpurun_out/r2_time3.log
