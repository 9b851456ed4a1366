set -x
timeout 600 python tools/r2_check.py check > gpurun_out/r2_check1.log 2>&1; echo "check rc=$?" >> gpurun_out/r2_check1.log
tail -30 gpurun_out/r2_check1.log
( timeout 200 python tools/r2_check.py time bilinear bicubic
for nc in 4 5 8; do PARADIS_SL_ROWS_NC=$nc timeout 200 python tools/r2_check.py time bilinear; done
for w in w20 w28 w32; do PARADIS_SL_LIB=build/variants/lib_$w.so timeout 200 python tools/r2_check.py time bilinear; done
PARADIS_SL_BWD=1 timeout 200 python tools/r2_check.py time bilinear ) > gpurun_out/r2_time1.log 2>&1
grep TIME gpurun_out/r2_time1.log
