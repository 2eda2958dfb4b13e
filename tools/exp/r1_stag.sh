#!/bin/bash
mkdir -p gpurun_out
for s in 0 2000 4000 6000 8000 12000; do
  MVRL_STAGGER_NS=$s python bench.py --steps 500 --warmup 20 --no-cpu > gpurun_out/r1_stag_$s.json 2>> gpurun_out/r1_stag.err
done
MVRL_STAGGER_NS=6000 python bench.py --steps 300 --warmup 20 --no-cpu --action-mode setpoint > gpurun_out/r1_stag_sp_6000.json 2>> gpurun_out/r1_stag.err
MVRL_STAGGER_NS=15000 python bench.py --steps 300 --warmup 20 --no-cpu --action-mode setpoint > gpurun_out/r1_stag_sp_15000.json 2>> gpurun_out/r1_stag.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r1_stag_*.json')):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], 'value %.4e'%d['value'], 'ms %.4f'%d['ms_per_step'])
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r1_stag.err
