#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rov6_gpu.py -m gpu -x -q -k "step_host" 2>&1 | tail -2
python bench.py --steps 100 --warmup 10 --no-cpu > gpurun_out/r1p_default.json 2>> gpurun_out/r1p.err
for r in "0.03,1.3" "0.05,1.3" "0.04,1.5" "0.06,1.25" "0.08,1.15"; do
  MVRL_HOST_RAMP=$r python bench.py --steps 100 --warmup 10 --no-cpu > gpurun_out/r1p_ramp_$r.json 2>> gpurun_out/r1p.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r1p_*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], 'value %.4e' % d['value'], 'ms %.4f' % d['ms_per_step'], 'e2e %.4e' % d['e2e']['value'], d['e2e']['chunks'])
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -5 gpurun_out/r1p.err
