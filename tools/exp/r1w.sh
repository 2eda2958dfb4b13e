#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r1w_tests.log 2>&1
tail -3 gpurun_out/r1w_tests.log
python bench.py --steps 500 --warmup 20 --no-cpu > gpurun_out/r1w_rpm.json 2>> gpurun_out/r1w.err
python bench.py --steps 300 --warmup 20 --no-cpu --action-mode setpoint > gpurun_out/r1w_sp.json 2>> gpurun_out/r1w.err
python bench.py --steps 300 --warmup 20 --no-cpu --action-mode force > gpurun_out/r1w_force.json 2>> gpurun_out/r1w.err
python bench.py --steps 300 --warmup 20 --no-cpu --dtype f64 > gpurun_out/r1w_f64.json 2>> gpurun_out/r1w.err
MVRL_NO_X2=1 python bench.py --steps 300 --warmup 20 --no-cpu > gpurun_out/r1w_scalar.json 2>> gpurun_out/r1w.err
python bench.py --workload rov3 --steps 300 --warmup 20 > gpurun_out/r1w_rov3.json 2>> gpurun_out/r1w.err
python bench.py --workload rov3 --steps 300 --warmup 20 --action-mode setpoint > gpurun_out/r1w_rov3_sp.json 2>> gpurun_out/r1w.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r1w_*.json')):
    try:
        d = json.load(open(f)); print(f.split('/')[-1], 'value %.4e' % d['value'], 'ms %.4f' % d['ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/r1w.err
