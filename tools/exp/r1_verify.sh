#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r1v_tests.log 2>&1
head -2 gpurun_out/r1v_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r1v_bench.json 2> gpurun_out/r1v_bench.err; echo rc=$? lines=$(wc -l < gpurun_out/r1v_bench.json)
python bench.py --workload auv --steps 500 --warmup 10 > gpurun_out/r1v_auv.json 2>> gpurun_out/r1v_bench.err
python - <<'PY'
import json
for f in ('r1v_bench','r1v_auv'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f,'value %.4e'%d['value'],'e2e',d.get('e2e',{}).get('value') if d.get('e2e') else None, 'frac', d['roofline']['frac'])
PY
