#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rov6_gpu.py tests/test_full_size_gpu.py -m gpu -x -q 2>&1 | tail -2
MVRL_PERSIST=1 MVRL_STAGGER_NS=1700 timeout 600 python -m pytest tests/test_rov6_gpu.py tests/test_full_size_gpu.py tests/test_awkward_sizes_gpu.py -m gpu -x -q 2>&1 | tail -2
MVRL_PERSIST=0 python bench.py --steps 500 --warmup 20 --no-cpu > gpurun_out/r1_pst_plain.json 2>> gpurun_out/r1_pst.err
for s in 0 800 1700 2500 3400; do
  MVRL_PERSIST=1 MVRL_STAGGER_NS=$s python bench.py --steps 500 --warmup 20 --no-cpu > gpurun_out/r1_pst_p1_$s.json 2>> gpurun_out/r1_pst.err
done
MVRL_PERSIST=1 MVRL_STAGGER_NS=4000 python bench.py --steps 300 --warmup 20 --no-cpu --action-mode setpoint > gpurun_out/r1_pst_sp_4000.json 2>> gpurun_out/r1_pst.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r1_pst_*.json')):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], 'value %.4e'%d['value'], 'ms %.4f'%d['ms_per_step'])
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r1_pst.err
