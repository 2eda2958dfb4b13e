#!/bin/bash
# round-1 session-3 experiment batch (run under gpurun); outputs into gpurun_out/
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r1k_tests.log 2>&1
tail -3 gpurun_out/r1k_tests.log
python bench.py > gpurun_out/r1k_bench.json 2> gpurun_out/r1k_bench.err
python - <<'PY'
import json; d=json.load(open('gpurun_out/r1k_bench.json')); print('default', '%.4e'%d['value'], 'e2e %.4e'%d['e2e']['value'])
PY
for c in 2 8 16 32; do
  python bench.py --steps 200 --warmup 20 --no-cpu --e2e-chunks $c > gpurun_out/r1k_e2e_c$c.json 2>> gpurun_out/r1k_bench.err
done
for n in 1022976 1136640 2097152; do
  python bench.py --steps 500 --warmup 20 --no-cpu --envs $n > gpurun_out/r1k_envs_$n.json 2>> gpurun_out/r1k_bench.err
done
python bench.py --steps 500 --warmup 20 --no-cpu --graph 0 > gpurun_out/r1k_nograph.json 2>> gpurun_out/r1k_bench.err
for b in 64 96 192 384; do
  MVRL_LIB=$PWD/marinevehiclereinforcementlearning_b200/libmvrl_b$b.so python bench.py --steps 500 --warmup 20 --no-cpu > gpurun_out/r1k_block_$b.json 2>> gpurun_out/r1k_bench.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r1k_*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], 'value %.4e' % d['value'], 'ms %.4f' % d['ms_per_step'], 'e2e %.4e' % d['e2e']['value'], 'chunks', d['e2e'].get('chunks'))
    except Exception as e:
        print(f, 'ERR', e)
PY
