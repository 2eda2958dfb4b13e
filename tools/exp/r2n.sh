#!/bin/bash
# round 2, GPU call N: TMA-fed persistent auv_step (MVRL_AUV_TMA=1) against the plain kernel
O=gpurun_out/r2n; mkdir -p $O
timeout 600 python -m pytest tests/test_auv_gpu.py -q -rA -k "tma" > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
grep -E "passed|failed|rc=|^FAILED|^E  " $O/pytest.log | tail -12
A="python bench.py --workload auv --steps 500 --warmup 50"
$A > $O/auv_plain.json 2>> $O/err.log
MVRL_AUV_TMA=1 $A > $O/auv_tma.json 2>> $O/err.log
MVRL_AUV_TMA=1 $A --envs 1048576 > $O/auv_tma_1m.json 2>> $O/err.log
$A --envs 1048576 > $O/auv_plain_1m.json 2>> $O/err.log
for f in $O/*.json; do python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1].split('/')[-1], '%.4g'%d['value'], '%.2f us'%(d['ms_per_step']*1e3))
PY
done
tail -3 $O/err.log
