#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:auv_step_kernel -s 450 -c 1 -o gpurun_out/prof_r1t_auv_steady -f python bench.py --workload auv --steps 500 --warmup 3 --graph 0 > gpurun_out/r1t_ncu.log 2>&1
tail -3 gpurun_out/r1t_ncu.log
