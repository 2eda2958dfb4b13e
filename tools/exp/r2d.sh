#!/bin/bash
# round 2, GPU call D: pipelined auv_step kernel, out-of-line trig / packed PID gains; same-box A/B against the round-1 library
O=gpurun_out/r2d; mkdir -p $O
P=$PWD/marinevehiclereinforcementlearning_b200
timeout 900 python -m pytest tests -m gpu -q -rA > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
grep -E "passed|failed|rc=" $O/pytest.log | tail -3
B="python bench.py --no-cpu --no-extra --steps 500 --warmup 50"
MVRL_LIB=$P/libmvrl_r1.so $B > $O/ab_r1_rpm.json 2>> $O/err.log
$B > $O/ab_r2_rpm.json 2>> $O/err.log
MVRL_LIB=$P/libmvrl_r1.so $B --action-mode setpoint > $O/ab_r1_sp.json 2>> $O/err.log
$B --action-mode setpoint > $O/ab_r2_sp.json 2>> $O/err.log
$B --action-mode force > $O/ab_r2_force.json 2>> $O/err.log
A="python bench.py --workload auv --steps 500 --warmup 50"
MVRL_LIB=$P/libmvrl_r1.so $A > $O/auv_r1.json 2>> $O/err.log
MVRL_AUV_NO_PIPELINE=1 $A > $O/auv_plain.json 2>> $O/err.log
$A > $O/auv_pipelined.json 2>> $O/err.log
$A --envs 1048576 > $O/auv_pipelined_1m.json 2>> $O/err.log
MVRL_AUV_NO_PIPELINE=1 $A --envs 1048576 > $O/auv_plain_1m.json 2>> $O/err.log
ncu --set full --clock-control none --import-source on -k regex:auv_step --launch-skip 280 -c 1 -o $O/auv python bench.py --workload auv --steps 20 --warmup 270 --graph 0 > $O/ncu_auv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rov6_step_kernel --launch-skip 8 -c 1 -o $O/rov6_sp python bench.py --no-cpu --no-extra --steps 5 --warmup 5 --graph 0 --action-mode setpoint > $O/ncu_sp.log 2>&1
ncu --set full --clock-control none -k regex:policy_act --launch-skip 3 -c 1 -o $O/policy python bench.py --workload rollout --steps 256 --warmup 128 > $O/ncu_policy.log 2>&1
ls $O
