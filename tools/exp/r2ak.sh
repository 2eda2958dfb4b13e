#!/bin/bash
# round 2, GPU call AK: vec_tools.EnvBlocks against whole-batch launches (bitwise), the bench contract with stream groups,
# and the default bench line with the calibrated choice of stream groups
O=gpurun_out/r2ak; mkdir -p $O
timeout 150 python -m pytest tests/test_vec_tools_gpu.py tests/test_bench_gpu.py -x -q --durations=5 > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -12 $O/pytest.log
timeout 150 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -3 $O/bench_default.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2ak/bench_default.json').read().strip().splitlines()[-1])
    print('headline %.4g' % d['value'], d['ms_per_step'], d['config']['stream_groups'], d['config']['stream_groups_tried_ms_per_step'], d.get('single_launch_per_step'))
    x = d['extra']
    print({k: (v['stream_groups'], '%.4g' % v['value'], v['stream_groups_tried_ms_per_step']) for k, v in x['single_gpu_shards_rpm_f32'].items()})
    print('setpoint %.4g' % x['setpoint_f32']['value'], x['setpoint_f32']['stream_groups'], 'force %.4g' % x['force_f32']['value'], x['force_f32']['stream_groups'])
    print('rollout %.4g' % x['config5_rollout']['value'], x['config5_rollout']['two_stream_groups'])
    print('e2e %.4g' % d['e2e']['value'], d['clocks'])
except Exception as e:
    print('failed', e)
PY
