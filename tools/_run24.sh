python tools/sanitize_case.py 2>&1 | tail -5
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_case.py > gpurun_out/racecheck_r2_full.log 2>&1; grep -c "=========" gpurun_out/racecheck_r2_full.log; head -c 6000 gpurun_out/racecheck_r2_full.log
