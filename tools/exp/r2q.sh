#!/bin/bash
# round 2, GPU call Q: actor kernel - observations prefetched one tile ahead; warp start offsets
O=gpurun_out/r2q; mkdir -p $O
P=$PWD/marinevehiclereinforcementlearning_b200
timeout 600 python -m pytest tests/test_policy_gpu.py -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log; tail -2 $O/pytest.log
R="python bench.py --workload rollout --steps 20 --warmup 3"
for v in _pf0 "" _st400 _st800; do
  MVRL_LIB=$P/libmvrl$v.so $R > $O/rollout$v.json 2>> $O/err.log
done
for f in $O/rollout*.json; do python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1].split('/')[-1], '%.4g'%d['value'], '%.2f us'%(d['ms_per_step']*1e3), {k:v for k,v in d.items() if 'us' in k or 'share' in k})
PY
done
tail -3 $O/err.log
