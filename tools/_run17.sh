timeout 300 python tools/r2_check.py check > gpurun_out/r2_check17.log 2>&1; grep -E "BAD|CHECK|WATCHDOG|Error" gpurun_out/r2_check17.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest17.log 2>&1; tail -12 gpurun_out/pytest17.log
grep -E "^config 1|^config 4|losses" gpurun_out/pytest17.log | head; cat gpurun_out/inductor_comparator.json
