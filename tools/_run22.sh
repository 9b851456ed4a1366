for n in 2 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2959$n bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/bench_r2_n$n.json 2> gpurun_out/bench_r2_n$n.err; echo "N=$n rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_n$n.json')); print($n, 'ms/step', round(d['ms_per_step'],4), 'phases', d['roofline']['phases_ms'], 'e2e ms', d.get('e2e',{}).get('ms_per_step'), 'batch ms', d.get('batch_sharded',{}).get('ms_per_step'))"
done
