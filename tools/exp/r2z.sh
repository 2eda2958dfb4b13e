#!/bin/bash
# round 2, GPU call Z: 64-thread CTAs at 144 registers (14 warps per SM, one wave at 131 072 envs) for the set-point / force kernels
O=gpurun_out/r2z; mkdir -p $O
P=$PWD/marinevehiclereinforcementlearning_b200
B="python bench.py --no-cpu --no-extra --steps 200 --warmup 20"
for v in "" _v5; do
  for m in setpoint force; do
    MVRL_LIB=$P/libmvrl$v.so $B --action-mode $m --envs 131072 > $O/${m}_128k$v.json 2>> $O/err.log
    MVRL_LIB=$P/libmvrl$v.so $B --action-mode $m > $O/${m}_1m$v.json 2>> $O/err.log
  done
  MVRL_LIB=$P/libmvrl$v.so python bench.py --workload rollout --steps 20 --warmup 3 > $O/rollout$v.json 2>> $O/err.log
done
for f in $O/*.json; do python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1].split('/')[-1], '%.4g'%d['value'], '%.2f us'%(d['ms_per_step']*1e3))
PY
done
tail -3 $O/err.log
