#!/bin/bash
# final round-1 evidence: tests, smoke, bench (ours + reference arm), launch list, ncu --set full of the headline kernel
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r1f_tests.log 2>&1
head -2 gpurun_out/r1f_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1f_smoke.log 2>&1; tail -1 gpurun_out/r1f_smoke.log
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r1f_bench_ref.json 2> gpurun_out/r1f_bench.err
python bench.py > gpurun_out/r1f_bench.json 2>> gpurun_out/r1f_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>> gpurun_out/r1f_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1f_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r1f_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rov6_step_kernel -s 5 -c 1 -o gpurun_out/prof_r1f_rov6 -f python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r1f_ncu1.log 2>&1
python bench.py --workload auv --steps 500 --warmup 10 > gpurun_out/r1f_auv.json 2>> gpurun_out/r1f_bench.err
python bench.py --workload rov3 --steps 300 --warmup 20 > gpurun_out/r1f_rov3.json 2>> gpurun_out/r1f_bench.err
python bench.py --workload rov3 --steps 300 --warmup 20 --action-mode setpoint > gpurun_out/r1f_rov3_sp.json 2>> gpurun_out/r1f_bench.err
python bench.py --workload rollout --steps 512 --warmup 128 > gpurun_out/r1f_rollout.json 2>> gpurun_out/r1f_bench.err
python bench.py --steps 300 --warmup 20 --no-cpu --action-mode setpoint > gpurun_out/r1f_sp.json 2>> gpurun_out/r1f_bench.err
python bench.py --steps 300 --warmup 20 --no-cpu --action-mode force > gpurun_out/r1f_force.json 2>> gpurun_out/r1f_bench.err
python bench.py --steps 300 --warmup 20 --no-cpu --dtype f64 > gpurun_out/r1f_f64.json 2>> gpurun_out/r1f_bench.err
python bench.py --steps 300 --warmup 20 --no-cpu --dtype f64 --envs 4096 > gpurun_out/r1f_f64_4096.json 2>> gpurun_out/r1f_bench.err
python bench.py --steps 500 --warmup 20 --no-cpu --n-sub 4 > gpurun_out/r1f_ns4.json 2>> gpurun_out/r1f_bench.err
python bench.py --steps 500 --warmup 20 --no-cpu --fast-math 1 > gpurun_out/r1f_fast.json 2>> gpurun_out/r1f_bench.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r1f_*.json')):
    try:
        d = json.load(open(f)); print(f.split('/')[-1], 'value %.4e' % d['value'], 'ms %.4f' % d['ms_per_step'], 'e2e', d.get('e2e',{}).get('value'), 'frac', d.get('roofline',{}).get('frac') if d.get('roofline') else None)
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/r1f_bench.err
