#!/bin/bash
# round 2, GPU call AH: programmatic dependent launch of the rollout's two kernels (MVRL_PDL=1)
O=gpurun_out/r2ah; mkdir -p $O
MVRL_PDL=1 timeout 120 python -m pytest tests/test_policy_gpu.py tests/test_dropin6_gpu.py -q -x > $O/pytest_pdl.log 2>&1; echo "pytest_pdl rc=$?"; tail -1 $O/pytest_pdl.log
R="timeout 120 python bench.py --workload rollout --steps 20 --warmup 3"
for v in 0 1 0 1; do MVRL_PDL=$v $R > $O/rollout_pdl${v}_$RANDOM.json 2>> $O/err.log; done
MVRL_PDL=1 timeout 120 python bench.py --no-cpu --no-extra --steps 200 --warmup 20 > $O/rpm_pdl1.json 2>> $O/err.log
timeout 120 python bench.py --no-cpu --no-extra --steps 200 --warmup 20 > $O/rpm_pdl0.json 2>> $O/err.log
for f in $O/r*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], '%.4g'%d['value'], '%.2f us'%(d['ms_per_step']*1e3), d.get('policy_and_bookkeeping_us_per_step'))
except Exception as e: print(sys.argv[1], 'failed')
PY
done
tail -3 $O/err.log
