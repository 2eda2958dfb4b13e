#!/bin/bash
# round 2, GPU call AF: double-buffered tcgen05.ld under the turn schedule
O=gpurun_out/r2af; mkdir -p $O
P=$PWD/marinevehiclereinforcementlearning_b200
timeout 60 python -m pytest tests/test_policy_gpu.py -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -1 $O/pytest.log
R="timeout 90 python bench.py --workload rollout --steps 20 --warmup 3"
for v in "" _np "" _np; do MVRL_LIB=$P/libmvrl$v.so $R > $O/rollout${v}_$RANDOM.json 2>> $O/err.log; done
for f in $O/rollout*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], '%.4g'%d['value'], '%.2f us'%(d['ms_per_step']*1e3), 'policy %.2f us'%d['policy_and_bookkeeping_us_per_step'])
except Exception as e: print(sys.argv[1], 'failed')
PY
done
tail -3 $O/err.log
