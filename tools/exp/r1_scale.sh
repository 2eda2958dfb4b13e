#!/bin/bash
mkdir -p gpurun_out
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 1000 --warmup 50 --no-cpu > gpurun_out/r1f_scale_$n.json 2> gpurun_out/r1f_scale_$n.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r1f_scale_$n.json')); print($n, 'value %.4e'%d['value'], 'ms %.4f'%d['ms_per_step'], 'e2e %.4e'%d['e2e']['value'])
except Exception as e: print($n, 'ERR', e)
PY
done
tail -3 gpurun_out/r1f_scale_8.err
