#!/bin/bash
# round 2, GPU call AD: the driver's sequence on the final tree + final bench lines, launch list and ncu summaries
O=gpurun_out/r2ad; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -3 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 300 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv python bench.py --no-cpu --steps 5 --warmup 3 --extra-steps 5 --rollout-len 8 > $O/ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:rov6_step_kernel --launch-skip 8 -c 1 -o $O/rov6_step python bench.py --no-cpu --no-extra --steps 5 --warmup 5 --graph 0 > $O/ncu_step.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:policy_act -c 1 --launch-skip 40 -o $O/actor python bench.py --workload rollout --steps 3 --warmup 1 --rollout-len 32 > $O/ncu_actor.log 2>&1
ls $O
