#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do python bench.py --workload auv --steps 500 --warmup 10 > gpurun_out/r1_auv3_$i.json 2>> gpurun_out/r1_auv3.err; done
python bench.py --workload auv --steps 2000 --warmup 50 > gpurun_out/r1_auv3_long.json 2>> gpurun_out/r1_auv3.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r1_auv3_*.json')):
    d=json.load(open(f)); print(f.split('/')[-1], 'value %.4e'%d['value'], 'ms %.4f'%d['ms_per_step'], d['episode_stats']['mean_length'])
PY
