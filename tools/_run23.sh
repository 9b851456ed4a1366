timeout 300 python bench.py --workload c2 --no-cpu --no-e2e > gpurun_out/bench_r2_c2.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_c2.json')); print('c2 ms/step', d['ms_per_step'], d['roofline']['phases_ms'])"
( echo "# compute-sanitizer, round 2 (tools/sanitize_case.py: library-chosen backward + forced row sweep, pad, dwconv, avgpool)"
echo "## memcheck"; timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_case.py 2>&1 | grep -E "ERROR SUMMARY|Invalid|case|dwconv|error" | head -20
echo "## racecheck"; timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_case.py 2>&1 | grep -E "RACECHECK SUMMARY|hazard|Race reported|case|dwconv" | sort | uniq -c | sort -rn | head -20 ) > gpurun_out/sanitizer_r2.txt 2>&1; cat gpurun_out/sanitizer_r2.txt
