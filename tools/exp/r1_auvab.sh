#!/bin/bash
mkdir -p gpurun_out
P=$PWD/marinevehiclereinforcementlearning_b200
for v in "" _pyb "" _pyb; do
  MVRL_LIB=$P/libmvrl$v.so python bench.py --workload auv --steps 500 --warmup 10 2>> gpurun_out/r1_auvab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'value %.4e'%d['value'], 'ms %.4f'%d['ms_per_step'])"
done
MVRL_LIB=$P/libmvrl_pyb.so python bench.py --steps 300 --warmup 20 --no-cpu 2>> gpurun_out/r1_auvab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rov6 pyb', 'value %.4e'%d['value'])"
python bench.py --steps 300 --warmup 20 --no-cpu 2>> gpurun_out/r1_auvab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rov6 new', 'value %.4e'%d['value'])"
