#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/exp/ws_check.py 2>&1 | tail -12
MVRL_WS=0 timeout 120 python bench.py --steps 300 --warmup 20 --no-cpu > gpurun_out/r1_ws0.json 2>> gpurun_out/r1_ws.err
MVRL_WS=1 timeout 120 python bench.py --steps 300 --warmup 20 --no-cpu > gpurun_out/r1_ws1.json 2>> gpurun_out/r1_ws.err
MVRL_WS=1 timeout 120 python bench.py --steps 300 --warmup 20 --no-cpu --n-sub 4 > gpurun_out/r1_ws1_ns4.json 2>> gpurun_out/r1_ws.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r1_ws*.json')):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], 'value %.4e'%d['value'], 'ms %.4f'%d['ms_per_step'], 'e2e %.4e'%d['e2e']['value'])
    except Exception as e: print(f,'ERR',e)
PY
tail -5 gpurun_out/r1_ws.err
