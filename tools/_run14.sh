timeout 300 python tools/r2_check.py check > gpurun_out/r2_check14.log 2>&1; echo "check rc=$?" >> gpurun_out/r2_check14.log
grep -E "BAD|CHECK|rc=|WATCHDOG|Error" gpurun_out/r2_check14.log
( timeout 100 python tools/r2_check.py time bilinear
PARADIS_SL_FWD=1 timeout 100 python tools/r2_check.py time bilinear ) > gpurun_out/r2_time14.log 2>&1
grep -E "TIME|WATCHDOG|Error" gpurun_out/r2_time14.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest14.log 2>&1; tail -25 gpurun_out/pytest14.log
