#!/bin/bash
mkdir -p gpurun_out
P=$PWD/marinevehiclereinforcementlearning_b200
for v in "" _b128m4 _b64m8 _b256m2 _b96m5; do
  MVRL_LIB=$P/libmvrl$v.so python bench.py --steps 500 --warmup 20 --no-cpu > gpurun_out/r1q_rpm$v.json 2>> gpurun_out/r1q.err
  MVRL_LIB=$P/libmvrl$v.so python bench.py --steps 300 --warmup 20 --no-cpu --action-mode setpoint > gpurun_out/r1q_sp$v.json 2>> gpurun_out/r1q.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r1q_*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], 'value %.4e' % d['value'], 'ms %.4f' % d['ms_per_step'], 'e2e %.4e' % d['e2e']['value'])
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -5 gpurun_out/r1q.err
