#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rov6_gpu.py -m gpu -x -q -k "warp_specialised or two_envs" 2>&1 | tail -2
for ns in 1 2; do for w in 0 1; do
MVRL_WS=$w timeout 120 python bench.py --steps 300 --warmup 20 --no-cpu --n-sub $ns > gpurun_out/r1_ws2_ns${ns}_w$w.json 2>> gpurun_out/r1_ws2.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r1_ws2_*.json')):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], 'value %.4e'%d['value'], 'ms %.4f'%d['ms_per_step'])
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r1_ws2.err
