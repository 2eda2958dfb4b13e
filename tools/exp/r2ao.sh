#!/bin/bash
# round 2, GPU call AO: the default bench line of the final tree (config 4 over shards on their own streams, 1 / 2 / 4 / 8 calibrated)
O=gpurun_out/r2ao; mkdir -p $O
timeout 55 python bench.py > $O/bench_default.json 2> $O/err.log; echo "bench rc=$?"; tail -2 $O/err.log
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2ao/bench_default.json').read().strip().splitlines()[-1])
    a = d['extra']['config4_auv_262144_envs']
    print('headline %.4g' % d['value'], d['config']['stream_groups'], 'auv %.4g' % a['value'], a['stream_groups'], a['stream_groups_tried_ms_per_step'], a['frac_of_hbm_peak'])
except Exception as e:
    print('failed', e)
PY
