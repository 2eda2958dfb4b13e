#!/bin/bash
# round 2, GPU call O: ncu of the TMA-fed auv_step
O=gpurun_out/r2o; mkdir -p $O
MVRL_AUV_TMA=1 ncu --set full --clock-control none --import-source on -k regex:auv_step --launch-skip 280 -c 1 -o $O/auv_tma python bench.py --workload auv --steps 20 --warmup 270 --graph 0 > $O/ncu.log 2>&1
tail -3 $O/ncu.log; ls $O
