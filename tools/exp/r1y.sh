#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r1y_tests.log 2>&1
head -3 gpurun_out/r1y_tests.log
python bench.py --steps 300 --warmup 20 --no-cpu --action-mode setpoint > gpurun_out/r1y_sp.json 2>> gpurun_out/r1y.err
python bench.py --steps 300 --warmup 20 --no-cpu --action-mode setpoint --dtype f64 > gpurun_out/r1y_sp_f64.json 2>> gpurun_out/r1y.err
python bench.py --workload rollout --steps 512 --warmup 128 > gpurun_out/r1y_rollout.json 2>> gpurun_out/r1y.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r1y_*.json')):
    try:
        d = json.load(open(f)); print(f.split('/')[-1], 'value %.4e' % d['value'], 'ms %.4f' % d['ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/r1y.err
