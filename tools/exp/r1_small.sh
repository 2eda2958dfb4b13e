#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rov6_gpu.py tests/test_awkward_sizes_gpu.py -m gpu -x -q -k "step_host or awkward or empty" 2>&1 | tail -2
for n in 4096 65536 262144; do
python bench.py --steps 200 --warmup 20 --no-cpu --envs $n > gpurun_out/r1_small_$n.json 2>> gpurun_out/r1_small.err
python bench.py --steps 200 --warmup 20 --no-cpu --envs $n --e2e-chunks 8 > gpurun_out/r1_small_c8_$n.json 2>> gpurun_out/r1_small.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r1_small_*.json')):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], 'value %.4e'%d['value'], 'e2e %.4e'%d['e2e']['value'], d['e2e']['chunks'])
    except Exception as e: print(f,'ERR',e)
PY
