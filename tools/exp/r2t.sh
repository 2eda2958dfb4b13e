#!/bin/bash
# round 2, GPU call T: tcgen05 actor, two threads per environment row (8 warps per tile) vs one
O=gpurun_out/r2t; mkdir -p $O
P=$PWD/marinevehiclereinforcementlearning_b200
timeout 180 python -m pytest tests/test_policy_gpu.py -q -x > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log; tail -12 $O/pytest.log
R="timeout 300 python bench.py --workload rollout --steps 20 --warmup 3"
for v in "" _s1; do MVRL_LIB=$P/libmvrl$v.so $R > $O/rollout_tc5$v.json 2>> $O/err.log; done
for f in $O/rollout*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], '%.4g'%d['value'], '%.2f us'%(d['ms_per_step']*1e3), {k:v for k,v in d.items() if 'us' in k or 'share' in k})
except Exception as e: print(sys.argv[1], 'failed', e)
PY
done
tail -5 $O/err.log
