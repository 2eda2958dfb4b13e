#!/bin/bash
mkdir -p gpurun_out
MVRL_WS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:rov6_step_ws -s 5 -c 1 -o gpurun_out/prof_r1_ws -f python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r1_wsprof.log 2>&1
tail -2 gpurun_out/r1_wsprof.log
