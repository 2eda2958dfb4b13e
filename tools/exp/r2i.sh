#!/bin/bash
# round 2, GPU call I: the driver's sequence on the final tree + final bench lines and launch list
O=gpurun_out/r2i; mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -3 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_20.json 2> $O/bench_20.err; echo "bench20 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv python bench.py --no-cpu --steps 5 --warmup 3 --extra-steps 5 --rollout-len 8 > $O/ncu_list.log 2>&1
ncu --set full --clock-control none -k regex:rov6_step_kernel --launch-skip 8 -c 1 -o $O/rov6_force python bench.py --no-cpu --no-extra --steps 5 --warmup 5 --graph 0 --action-mode force > $O/ncu_force.log 2>&1
ls $O
