#!/bin/bash
# round 2, GPU call C: full GPU suite on the final kernels; same-box A/B against the round-1 library; bench line; ncu of the new auv kernel
O=gpurun_out/r2c; mkdir -p $O
P=$PWD/marinevehiclereinforcementlearning_b200
python -m pytest tests -m gpu -q -rA > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
grep -E "passed|failed|rc=" $O/pytest.log | tail -3
B="python bench.py --no-cpu --no-extra --steps 500 --warmup 50"
for rep in 1 2; do
  MVRL_LIB=$P/libmvrl_r1.so $B > $O/ab_r1_rpm_$rep.json 2>> $O/err.log
  $B > $O/ab_r2_rpm_$rep.json 2>> $O/err.log
done
MVRL_LIB=$P/libmvrl_r1.so $B --action-mode setpoint > $O/ab_r1_sp.json 2>> $O/err.log
$B --action-mode setpoint > $O/ab_r2_sp.json 2>> $O/err.log
MVRL_LIB=$P/libmvrl_r1.so $B --action-mode force > $O/ab_r1_force.json 2>> $O/err.log
$B --action-mode force > $O/ab_r2_force.json 2>> $O/err.log
MVRL_LIB=$P/libmvrl_r1.so python bench.py --workload auv --steps 500 --warmup 50 > $O/ab_r1_auv.json 2>> $O/err.log
python bench.py --workload auv --steps 500 --warmup 50 > $O/ab_r2_auv.json 2>> $O/err.log
python bench.py --steps 1000 --warmup 50 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2>> $O/err.log
ncu --set full --clock-control none --import-source on -k regex:auv_step_kernel --launch-skip 280 -c 1 -o $O/auv python bench.py --workload auv --steps 20 --warmup 270 --graph 0 > $O/ncu_auv.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --no-cpu --steps 5 --warmup 3 --extra-steps 5 > $O/ncu_list.log 2>&1
ls $O
