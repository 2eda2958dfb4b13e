#!/bin/bash
# round 2, multi-GPU call W: the final build on 8 GPUs (default bench line: weak headline + strong key + every extra leg incl. the
# rollout with the tcgen05 actor and its stats all-reduce), then the reference arm as the driver launches it
O=gpurun_out/r2w; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29708 bench.py --gpus 8 --steps 300 --warmup 20 > $O/bench_8.json 2>> $O/err.log; echo "bench 8 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29718 bench.py --gpus 8 --steps 50 --warmup 5 --workload rollout > $O/rollout_8.json 2>> $O/err.log; echo "rollout 8 rc=$?"
tail -5 $O/err.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w/bench_8.json').read().strip().splitlines()[-1])
print('weak', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
print('strong', d.get('strong'))
print('rollout', d['extra'].get('config5_rollout'))
PY
