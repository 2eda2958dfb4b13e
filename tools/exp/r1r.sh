#!/bin/bash
mkdir -p gpurun_out
P=$PWD/marinevehiclereinforcementlearning_b200
for v in "" _u1 _u2; do
  MVRL_LIB=$P/libmvrl$v.so python bench.py --steps 300 --warmup 20 --no-cpu --action-mode setpoint > gpurun_out/r1r_sp$v.json 2>> gpurun_out/r1r.err
  MVRL_LIB=$P/libmvrl$v.so python bench.py --steps 300 --warmup 20 --no-cpu --action-mode force > gpurun_out/r1r_force$v.json 2>> gpurun_out/r1r.err
  MVRL_LIB=$P/libmvrl$v.so python bench.py --steps 300 --warmup 20 --no-cpu > gpurun_out/r1r_rpm$v.json 2>> gpurun_out/r1r.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r1r_*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], 'value %.4e' % d['value'], 'ms %.4f' % d['ms_per_step'])
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -5 gpurun_out/r1r.err
