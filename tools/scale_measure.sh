#!/bin/bash
# Strong-scaling lines of bench.py on one multi-GPU box (run under `gpurun --gpus 8` from the repo root).
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2959$n bench.py --gpus $n --steps 50 --warmup 5 2> gpurun_out/bench_r2_n$n.err | grep '^{' > gpurun_out/bench_r2_latband$n.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_r2_latband$n.json')); print($n, 'ms/step', round(d['ms_per_step'],4), 'phases', {k:round(v,4) for k,v in d['roofline']['phases_ms'].items()}, 'e2e ms', round(d['e2e']['ms_per_step'],2), 'batch ms', round(d['batch_sharded']['ms_per_step'],4))"
done
