timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest16.log 2>&1; tail -30 gpurun_out/pytest16.log
grep -E "config 1|config 4|c3 smooth|c3 exact" gpurun_out/pytest16.log | head -20
