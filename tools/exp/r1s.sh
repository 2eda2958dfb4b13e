#!/bin/bash
mkdir -p gpurun_out
python bench.py --workload auv --steps 200 --warmup 10 > gpurun_out/r1s_auv.json 2> gpurun_out/r1s.err
python bench.py --workload auv --steps 200 --warmup 10 --graph 0 > gpurun_out/r1s_auv_nograph.json 2>> gpurun_out/r1s.err
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r1s_short.json 2>> gpurun_out/r1s.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1s.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r1s_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rov6_step_kernel -s 5 -c 1 -o gpurun_out/prof_r1s_rov6 -f python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r1s_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:auv_step_kernel -s 5 -c 1 -o gpurun_out/prof_r1s_auv -f python bench.py --workload auv --steps 20 --warmup 3 --graph 0 > gpurun_out/r1s_ncu2.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r1s_*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], 'value %.4e' % d['value'], 'ms %.4f' % d['ms_per_step'], d.get('roofline',{}).get('frac'))
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -5 gpurun_out/r1s.err
