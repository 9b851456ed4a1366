#!/bin/bash
# Round-end measurement on one B200 (run under gpurun from the repo root): tests, smoke, bench lines, launch list with
# DRAM bytes, ncu summary of the two dominant kernels.  Outputs under gpurun_out/.
timeout 1700 python -m pytest tests -q -m gpu > gpurun_out/pytest_final.log 2>&1; tail -4 gpurun_out/pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; cut -c1-400 gpurun_out/bench_r2.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference.json 2>/dev/null
timeout 600 python bench.py --interp bicubic --no-cpu --no-e2e > gpurun_out/bench_r2_bicubic.json 2>/dev/null
timeout 600 python bench.py --workload c2 --no-cpu --no-e2e > gpurun_out/bench_r2_c2.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python tools/prof_step.py 64 bilinear fast 2 6.0 > /dev/null 2>&1
python tools/ncu_traffic.py gpurun_out/launches_r2.csv gpurun_out/traffic_r2.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sl_fwd_kernel|sl_bwd_rows_kernel" -c 2 -f -o gpurun_out/prof_r2_final python tools/prof_step.py 64 bilinear fast 1 6.0 > gpurun_out/ncu_r2_final.log 2>&1; tail -1 gpurun_out/ncu_r2_final.log
