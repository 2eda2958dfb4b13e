#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 300 --warmup 20 --no-cpu > gpurun_out/r1g_scale_2.json 2> gpurun_out/r1g_scale_2.err
echo rc=$?
wc -c gpurun_out/r1g_scale_2.json
tail -20 gpurun_out/r1g_scale_2.err
cat gpurun_out/r1g_scale_2.json | head -c 600
