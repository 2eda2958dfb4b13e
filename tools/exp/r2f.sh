#!/bin/bash
# round 2, GPU call F: validation of the shipped build - GPU suite, smoke, same-box A/B, bench lines, ncu of the shipped kernels
O=gpurun_out/r2f; mkdir -p $O
P=$PWD/marinevehiclereinforcementlearning_b200
timeout 900 python -m pytest tests -m gpu -q -rA > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
grep -E "passed|failed|rc=" $O/pytest.log | tail -3
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
B="python bench.py --no-cpu --no-extra --steps 500 --warmup 50"
MVRL_LIB=$P/libmvrl_r1.so $B > $O/ab_r1_rpm.json 2>> $O/err.log
$B > $O/ab_r2_rpm.json 2>> $O/err.log
MVRL_LIB=$P/libmvrl_r1.so $B --action-mode setpoint > $O/ab_r1_sp.json 2>> $O/err.log
$B --action-mode setpoint > $O/ab_r2_sp.json 2>> $O/err.log
MVRL_LIB=$P/libmvrl_r1.so $B --action-mode setpoint --dtype f64 --envs 262144 > $O/ab_r1_sp64.json 2>> $O/err.log
$B --action-mode setpoint --dtype f64 --envs 262144 > $O/ab_r2_sp64.json 2>> $O/err.log
MVRL_LIB=$P/libmvrl_r1.so $B --action-mode force > $O/ab_r1_force.json 2>> $O/err.log
$B --action-mode force > $O/ab_r2_force.json 2>> $O/err.log
MVRL_LIB=$P/libmvrl_r1.so python bench.py --workload auv --steps 500 --warmup 50 > $O/ab_r1_auv.json 2>> $O/err.log
python bench.py --workload auv --steps 500 --warmup 50 > $O/ab_r2_auv.json 2>> $O/err.log
python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2>> $O/err.log
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:rov6_step_kernel --launch-skip 8 -c 1 -o $O/rov6_rpm python bench.py --no-cpu --no-extra --steps 5 --warmup 5 --graph 0 > $O/ncu_rpm.log 2>&1
$NCU -k regex:rov6_step_kernel --launch-skip 8 -c 1 -o $O/rov6_sp python bench.py --no-cpu --no-extra --steps 5 --warmup 5 --graph 0 --action-mode setpoint > $O/ncu_sp.log 2>&1
$NCU -k regex:auv_step --launch-skip 280 -c 1 -o $O/auv python bench.py --workload auv --steps 20 --warmup 270 --graph 0 > $O/ncu_auv.log 2>&1
$NCU -k regex:policy_act --launch-skip 3 -c 1 -o $O/policy python bench.py --workload rollout --steps 256 --warmup 128 > $O/ncu_policy.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv python bench.py --no-cpu --steps 5 --warmup 3 --extra-steps 5 --rollout-len 8 > $O/ncu_list.log 2>&1
ls $O
