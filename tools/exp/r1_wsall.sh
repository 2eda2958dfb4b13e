#!/bin/bash
MVRL_WS=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 1 --steps 100 --warmup 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('torchrun n=1', 'value %.4e'%d['value'], 'cpu', d['cpu_baseline'] is not None)"
