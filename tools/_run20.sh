timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest20.log 2>&1; tail -6 gpurun_out/pytest20.log; grep -E "^config|c3 smooth.*exact 6.0|c3 exact" gpurun_out/pytest20.log | cut -c1-330 | head -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; cut -c1-900 gpurun_out/bench_r2.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r2_reference.json
timeout 600 python bench.py --interp bicubic --no-cpu --no-e2e > gpurun_out/bench_r2_bicubic.json 2>/dev/null; cut -c1-400 gpurun_out/bench_r2_bicubic.json
timeout 600 python bench.py --workload c2 --no-cpu --no-e2e > gpurun_out/bench_r2_c2.json 2>/dev/null; cut -c1-400 gpurun_out/bench_r2_c2.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python tools/prof_step.py 64 bilinear fast 2 6.0 > /dev/null 2>&1
python tools/ncu_traffic.py gpurun_out/launches_r2.csv gpurun_out/traffic_r2.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sl_fwd_kernel|sl_bwd_rows_kernel" -c 2 -f -o gpurun_out/prof_r2_final python tools/prof_step.py 64 bilinear fast 1 6.0 > gpurun_out/ncu_r2_final.log 2>&1; tail -2 gpurun_out/ncu_r2_final.log
