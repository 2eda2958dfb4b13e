#!/bin/bash
# round 2, GPU call V: actor tests (both kernels), bench legs of the rollout with both actors, ncu summary
O=gpurun_out/r2v; mkdir -p $O
timeout 300 python -m pytest tests/test_policy_gpu.py -q -x > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log; tail -12 $O/pytest.log
R="timeout 300 python bench.py --workload rollout --steps 20 --warmup 3"
$R > $O/rollout_tcgen05.json 2>> $O/err.log
MVRL_POLICY_MMA_SYNC=1 $R > $O/rollout_mma_sync.json 2>> $O/err.log
for f in $O/rollout*.json; do python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1].split('/')[-1], '%.4g'%d['value'], '%.2f us'%(d['ms_per_step']*1e3), {k:v for k,v in d.items() if 'us' in k or 'share' in k})
PY
done
tail -3 $O/err.log
