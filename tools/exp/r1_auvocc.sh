#!/bin/bash
mkdir -p gpurun_out
P=$PWD/marinevehiclereinforcementlearning_b200
for v in "" _a128m6 _a128m8 _a64m12 _a256m3 _a64m16; do
  MVRL_LIB=$P/libmvrl$v.so python bench.py --workload auv --steps 500 --warmup 10 > gpurun_out/r1_auvocc$v.json 2>> gpurun_out/r1_auvocc.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r1_auvocc*.json')):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], 'value %.4e'%d['value'], 'ms %.4f'%d['ms_per_step'])
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r1_auvocc.err
