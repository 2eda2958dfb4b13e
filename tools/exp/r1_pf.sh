#!/bin/bash
mkdir -p gpurun_out
for t in 0 148 296 444 888; do
  MVRL_PREFETCH_TILES=$t python bench.py --steps 500 --warmup 20 --no-cpu > gpurun_out/r1_pf_$t.json 2>> gpurun_out/r1_pf.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r1_pf_*.json')):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], 'value %.4e'%d['value'], 'ms %.4f'%d['ms_per_step'])
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r1_pf.err
