#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_auv_gpu.py tests/test_vec_tools_gpu.py tests/test_awkward_sizes_gpu.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload auv --steps 500 --warmup 10 > gpurun_out/r1u_auv.json 2> gpurun_out/r1u.err
python bench.py --workload auv --steps 500 --warmup 10 --field noise > gpurun_out/r1u_auv_noise.json 2>> gpurun_out/r1u.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r1u_*.json')):
    d = json.load(open(f)); print(f.split('/')[-1], 'value %.4e' % d['value'], 'ms %.4f' % d['ms_per_step'], d.get('roofline',{}).get('frac'))
PY
tail -3 gpurun_out/r1u.err
