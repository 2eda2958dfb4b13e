#!/bin/bash
mkdir -p gpurun_out
for r in "0.03,1.3"; do
  MVRL_HOST_RAMP=$r python bench.py --steps 100 --warmup 10 --no-cpu > gpurun_out/r1m_ramp_$r.json 2>> gpurun_out/r1m.err
done
MVRL_PERSIST=1 ncu --set full --clock-control none --import-source on -k regex:rov6_step_kernel -s 5 -c 1 -o gpurun_out/prof_r1m_persist -f python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r1m_ncu.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r1m_*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], 'value %.4e' % d['value'], 'ms %.4f' % d['ms_per_step'], 'e2e %.4e' % d['e2e']['value'], d['e2e'].get('pieces'))
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -5 gpurun_out/r1m.err
