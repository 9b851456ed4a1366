timeout 300 python tools/r2_check.py check > gpurun_out/r2_check8.log 2>&1; echo "check rc=$?" >> gpurun_out/r2_check8.log
grep -E "BAD|CHECK|rc=|WATCHDOG|Error" gpurun_out/r2_check8.log
( timeout 100 python tools/r2_check.py time bilinear bicubic
for w0 in 0 12 24; do PARADIS_SL_ROWS_W0=$w0 timeout 100 python tools/r2_check.py time bilinear; done
for nc in 4 6 7; do PARADIS_SL_ROWS_NC=$nc timeout 100 python tools/r2_check.py time bilinear; done
for v in NOCONSUME NOPRODUCE; do PARADIS_SL_LIB=build/variants/lib_$v.so timeout 100 python tools/r2_check.py time bilinear; done ) > gpurun_out/r2_time8.log 2>&1
grep -E "TIME|WATCHDOG|Error" gpurun_out/r2_time8.log
