#!/bin/bash
# round 2, GPU call U: ncu of the tcgen05 actor
O=gpurun_out/r2u; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:policy_act -c 1 --launch-skip 40 -o $O/actor python bench.py --workload rollout --steps 3 --warmup 1 --rollout-len 32 > $O/ncu.log 2>&1
tail -3 $O/ncu.log; ls $O
