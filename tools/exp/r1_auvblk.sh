#!/bin/bash
mkdir -p gpurun_out
P=$PWD/marinevehiclereinforcementlearning_b200
for v in "" _ab64 _ab256; do
  MVRL_LIB=$P/libmvrl$v.so python bench.py --workload auv --steps 500 --warmup 10 2>> gpurun_out/r1_auvblk.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'value %.4e'%d['value'], 'ms %.4f'%d['ms_per_step'])"
done
