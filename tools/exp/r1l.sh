#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r1l_tests.log 2>&1
tail -3 gpurun_out/r1l_tests.log
for p in 0 1; do
  MVRL_PERSIST=$p python bench.py --steps 500 --warmup 20 --no-cpu > gpurun_out/r1l_persist$p.json 2>> gpurun_out/r1l.err
  MVRL_PERSIST=$p python bench.py --steps 300 --warmup 20 --no-cpu --action-mode setpoint > gpurun_out/r1l_sp_persist$p.json 2>> gpurun_out/r1l.err
  MVRL_PERSIST=$p python bench.py --steps 300 --warmup 20 --no-cpu --dtype f64 > gpurun_out/r1l_f64_persist$p.json 2>> gpurun_out/r1l.err
  MVRL_PERSIST=$p MVRL_NO_X2=1 python bench.py --steps 300 --warmup 20 --no-cpu > gpurun_out/r1l_scalar_persist$p.json 2>> gpurun_out/r1l.err
  MVRL_PERSIST=$p python bench.py --steps 500 --warmup 20 --no-cpu --n-sub 4 > gpurun_out/r1l_ns4_persist$p.json 2>> gpurun_out/r1l.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r1l_*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], 'value %.4e' % d['value'], 'ms %.4f' % d['ms_per_step'], 'e2e %.4e' % d['e2e']['value'])
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -5 gpurun_out/r1l.err
