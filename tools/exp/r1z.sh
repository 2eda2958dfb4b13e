#!/bin/bash
mkdir -p gpurun_out
P=$PWD/marinevehiclereinforcementlearning_b200
for v in "" _prev; do
MVRL_LIB=$P/libmvrl$v.so python bench.py --steps 200 --warmup 20 --no-cpu --action-mode setpoint --dtype f64 > gpurun_out/r1z_sp_f64$v.json 2>> gpurun_out/r1z.err
MVRL_LIB=$P/libmvrl$v.so python bench.py --steps 200 --warmup 20 --no-cpu --action-mode setpoint --dtype f64 --envs 262144 > gpurun_out/r1z_sp_f64_256k$v.json 2>> gpurun_out/r1z.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r1z_*.json')):
    try:
        d = json.load(open(f)); print(f.split('/')[-1], 'value %.4e' % d['value'], 'ms %.4f' % d['ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/r1z.err
