#!/bin/bash
# round 2, GPU call AL: the EnvBlocks test (fixed: episode_stats() resets on read) and the bench contract tests; launch list
O=gpurun_out/r2al; mkdir -p $O
timeout 150 python -m pytest tests/test_vec_tools_gpu.py tests/test_bench_gpu.py -x -q --durations=5 > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -12 $O/pytest.log
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches.csv python bench.py --no-cpu --no-extra --steps 5 --warmup 3 > $O/ncu_list.log 2>&1; echo "ncu rc=$?"
grep -c rov6_step_kernel $O/launches.csv
