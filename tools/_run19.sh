timeout 300 python tools/r2_check.py check > gpurun_out/r2_check19.log 2>&1; grep -E "BAD|CHECK|WATCHDOG|Error" gpurun_out/r2_check19.log
( timeout 100 python tools/r2_check.py time bilinear
PARADIS_SL_LIB=build/variants/lib_nopf.so timeout 100 python tools/r2_check.py time bilinear ) > gpurun_out/r2_time19.log 2>&1
grep -E "TIME|WATCHDOG|Error" gpurun_out/r2_time19.log
