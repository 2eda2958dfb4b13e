#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rov6_gpu.py tests/test_full_size_gpu.py tests/test_awkward_sizes_gpu.py tests/test_dropin6_gpu.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 500 --warmup 20 --no-cpu > gpurun_out/r1_walk_rpm.json 2>> gpurun_out/r1_walk.err
python bench.py --steps 500 --warmup 20 --no-cpu > gpurun_out/r1_walk_rpm2.json 2>> gpurun_out/r1_walk.err
python bench.py --steps 300 --warmup 20 --no-cpu --action-mode setpoint > gpurun_out/r1_walk_sp.json 2>> gpurun_out/r1_walk.err
python bench.py --steps 300 --warmup 20 --no-cpu --dtype f64 > gpurun_out/r1_walk_f64.json 2>> gpurun_out/r1_walk.err
MVRL_NO_X2=1 python bench.py --steps 300 --warmup 20 --no-cpu > gpurun_out/r1_walk_scalar.json 2>> gpurun_out/r1_walk.err
python bench.py --steps 500 --warmup 20 --no-cpu --n-sub 4 > gpurun_out/r1_walk_ns4.json 2>> gpurun_out/r1_walk.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r1_walk_*.json')):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], 'value %.4e'%d['value'], 'ms %.4f'%d['ms_per_step'])
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r1_walk.err
