#!/bin/bash
# round 2, GPU call K: launch shapes of the rpm step kernel with literal constants (64-thread CTAs at 144 / 128 registers =
# 14 / 16 warps per SM) at 1 Mi and 131 072 envs; auv_step x2 at 12 / 16 warps per SM
O=gpurun_out/r2k; mkdir -p $O
P=$PWD/marinevehiclereinforcementlearning_b200
B="python bench.py --no-cpu --no-extra --steps 300 --warmup 30"
for v in "" _v5 _v3; do
  L=$P/libmvrl$v.so
  MVRL_LIB=$L $B > $O/rpm_1m$v.json 2>> $O/err.log
  MVRL_LIB=$L $B --envs 131072 > $O/rpm_128k$v.json 2>> $O/err.log
  MVRL_LIB=$L $B --envs 262144 > $O/rpm_256k$v.json 2>> $O/err.log
  MVRL_LIB=$L $B --n-sub 4 > $O/rpm_1m_nsub4$v.json 2>> $O/err.log
  MVRL_LIB=$L python bench.py --workload auv --steps 500 --warmup 50 > $O/auv_x2$v.json 2>> $O/err.log
done
MVRL_AUV_NO_X2=1 python bench.py --workload auv --steps 500 --warmup 50 > $O/auv_one.json 2>> $O/err.log
for f in $O/*.json; do python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1].split('/')[-1], '%.4g'%d['value'], '%.2f us'%(d['ms_per_step']*1e3))
PY
done
tail -3 $O/err.log
