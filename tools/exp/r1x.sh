#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:rov6_step_kernel -s 5 -c 1 -o gpurun_out/prof_r1x_sp -f python bench.py --steps 10 --warmup 3 --no-cpu --action-mode setpoint > gpurun_out/r1x_ncu.log 2>&1
tail -2 gpurun_out/r1x_ncu.log
