timeout 400 python tools/r2_check.py check > gpurun_out/r2_check2.log 2>&1; echo "check rc=$?" >> gpurun_out/r2_check2.log
tail -25 gpurun_out/r2_check2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sl_bwd_rows -c 1 -f -o gpurun_out/prof_rows1 python tools/prof_step.py 16 bilinear fast 1 6.0 > gpurun_out/ncu_rows1.log 2>&1
tail -3 gpurun_out/ncu_rows1.log
