#!/bin/bash
# Strong-scaling lines of bench.py on one multi-GPU box (run under `gpurun --gpus 8` from the repo root).
# usage: [PSL_BENCH_FLAGS="--pull-all" PSL_TAG=_pull] scale_measure.sh [N ...]   (default 2 4 8)
for n in ${@:-2 4 8}; do
  out=gpurun_out/bench_r2_latband$n$PSL_TAG.json
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2959$n bench.py --gpus $n --steps 50 --warmup 5 $PSL_BENCH_FLAGS 2> gpurun_out/bench_r2_n$n$PSL_TAG.err | grep '^{' > $out
  python -c "
import json; d=json.load(open('$out')); print($n, '$PSL_TAG', 'ms/step', round(d['ms_per_step'],4), 'phases', {k:round(v,4) for k,v in d['roofline']['phases_ms'].items()}, 'e2e ms', round(d['e2e']['ms_per_step'],2), 'batch ms', round(d['batch_sharded']['ms_per_step'],4))" || tail -5 gpurun_out/bench_r2_n$n$PSL_TAG.err
done
