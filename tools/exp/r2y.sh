#!/bin/bash
# round 2, GPU call Y: e2e (host buffers) with a tapered piece plan: first and last piece smaller than the others
O=gpurun_out/r2y; mkdir -p $O
B="python bench.py --no-cpu --no-extra --steps 100 --warmup 10"
run() { # name, env, chunks
  env $2 $B --e2e-chunks $3 > $O/$1.json 2>> $O/err.log
  python - $O/$1.json $1 <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], 'e2e %.4g'%d['e2e']['value'], 'pieces', d['e2e']['chunks'])
PY
}
run eq8 MVRL_X=0 8
run eq8b MVRL_X=0 8
run t9_05 MVRL_HOST_TAPER=0.5 9
run t10_05 MVRL_HOST_TAPER=0.5 10
run t10_03 MVRL_HOST_TAPER=0.3 10
run t12_05 MVRL_HOST_TAPER=0.5 12
run t9_025 MVRL_HOST_TAPER=0.25 9
run eq10 MVRL_X=0 10
tail -3 $O/err.log
