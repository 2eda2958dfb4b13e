#!/bin/bash
# round 2, GPU call AN (last seconds of the budget): EnvShards test; legacy auv_step as shards on their own streams
O=gpurun_out/r2an; mkdir -p $O
timeout 40 python -m pytest tests/test_vec_tools_gpu.py -x -q -k "shards or blocks" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 40 python bench.py --workload auv --steps 200 --warmup 10 > $O/auv_auto.json 2> $O/err.log; echo "auv rc=$?"; tail -2 $O/err.log
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2an/auv_auto.json').read().strip().splitlines()[-1])
    print('auv %.4g' % d['value'], d['ms_per_step'], d['config']['stream_groups'], d['config']['stream_groups_tried_ms_per_step'], d['roofline']['frac'])
except Exception as e:
    print('failed', e)
PY
