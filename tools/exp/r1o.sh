#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rov6_gpu.py tests/test_awkward_sizes_gpu.py -m gpu -x -q 2>&1 | tail -3
for c in 4 6 8 12 16; do
  python bench.py --steps 100 --warmup 10 --no-cpu --e2e-chunks $c > gpurun_out/r1o_c$c.json 2>> gpurun_out/r1o.err
done
MVRL_HOST_NO_MAP=1 python bench.py --steps 100 --warmup 10 --no-cpu --e2e-chunks 4 > gpurun_out/r1o_nomap_c4.json 2>> gpurun_out/r1o.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r1o_*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], 'value %.4e' % d['value'], 'ms %.4f' % d['ms_per_step'], 'e2e %.4e' % d['e2e']['value'])
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -5 gpurun_out/r1o.err
