#!/bin/bash
# round 2, GPU call AG: refresh of the final single-GPU lines after the last actor changes
O=gpurun_out/r2ag; mkdir -p $O
timeout 300 python -m pytest tests/test_policy_gpu.py tests/test_bench_gpu.py -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -1 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv python bench.py --no-cpu --steps 5 --warmup 3 --extra-steps 5 --rollout-len 8 > $O/ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:policy_act -c 1 --launch-skip 40 -o $O/actor python bench.py --workload rollout --steps 3 --warmup 1 --rollout-len 32 > $O/ncu_actor.log 2>&1
ls $O
