timeout 300 python tools/r2_check.py check > gpurun_out/r2_check13.log 2>&1; echo "check rc=$?" >> gpurun_out/r2_check13.log
grep -E "BAD|CHECK|rc=|WATCHDOG|Error" gpurun_out/r2_check13.log
( timeout 100 python tools/r2_check.py time bilinear
PARADIS_SL_FWD=1 timeout 100 python tools/r2_check.py time bilinear ) > gpurun_out/r2_time13.log 2>&1
grep -E "TIME|WATCHDOG|Error" gpurun_out/r2_time13.log
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest13.log 2>&1; tail -15 gpurun_out/pytest13.log
