#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r1_topo.txt 2>&1
for f in /sys/bus/pci/devices/*/numa_node; do :; done
lscpu | grep -i -E "numa|socket|^CPU\(s\)" >> gpurun_out/r1_topo.txt
for numa in 0 1; do for n in 2 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n+10*numa)) bench.py --gpus $n --steps 200 --warmup 20 --no-cpu --numa $numa > gpurun_out/r1_numa${numa}_$n.json 2> gpurun_out/r1_numa${numa}_$n.err
done; done
python bench.py --steps 200 --warmup 20 --no-cpu --numa 1 > gpurun_out/r1_numa1_1.json 2> gpurun_out/r1_numa1_1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r1_numa*_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], 'value %.4e'%d['value'], 'e2e %.4e'%d['e2e']['value'], 'node', d['e2e'].get('host_numa_node'))
    except Exception as e: print(f,'ERR',e)
PY
cat gpurun_out/r1_topo.txt | head -30
