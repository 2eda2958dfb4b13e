#!/bin/bash
# round 2, GPU call AJ: independent blocks of environments as chains on their own streams (bench.py --stream-groups G):
# the 8-GPU shard size of the strong-scaling leg, and the rollout (one block's actor beside another block's env step)
O=gpurun_out/r2aj; mkdir -p $O
B="timeout 60 python bench.py --no-cpu --no-extra --steps 300 --warmup 20"
for g in 1 2 4; do $B --envs 131072 --stream-groups $g > $O/rpm_131072_g$g.json 2>> $O/err.log; done
$B --envs 262144 --stream-groups 2 > $O/rpm_262144_g2.json 2>> $O/err.log
$B --envs 65536 --stream-groups 2 > $O/rpm_65536_g2.json 2>> $O/err.log
$B --stream-groups 2 > $O/rpm_1048576_g2.json 2>> $O/err.log
R="timeout 60 python bench.py --workload rollout --steps 256 --warmup 3"
for g in 1 2 4; do $R --stream-groups $g > $O/rollout_g$g.json 2>> $O/err.log; done
for f in $O/*.json; do python - $f <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], '%.4g env-steps/s' % d['value'], '%.2f us' % (d['ms_per_step'] * 1e3), d.get('policy_and_bookkeeping_us_per_step'))
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
done
tail -5 $O/err.log
