#!/bin/bash
# round 2, GPU call H: fp32 step kernels with compile-time constants (CONSTP) against the same kernels reading them from the argument
O=gpurun_out/r2h; mkdir -p $O
P=$PWD/marinevehiclereinforcementlearning_b200
timeout 900 python -m pytest tests -m gpu -q -rA > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
grep -E "passed|failed|rc=" $O/pytest.log | tail -3
B="python bench.py --no-cpu --no-extra --steps 500 --warmup 50"
for m in rpm setpoint force; do
  MVRL_LIB=$P/libmvrl_r1.so $B --action-mode $m > $O/r1_$m.json 2>> $O/err.log
  MVRL_NO_CONSTP=1 $B --action-mode $m > $O/arg_$m.json 2>> $O/err.log
  $B --action-mode $m > $O/lit_$m.json 2>> $O/err.log
done
MVRL_NO_CONSTP=1 python bench.py --workload rollout --steps 512 --warmup 128 > $O/arg_rollout.json 2>> $O/err.log
python bench.py --workload rollout --steps 512 --warmup 128 > $O/lit_rollout.json 2>> $O/err.log
$B --envs 131072 > $O/lit_rpm_131072.json 2>> $O/err.log
ncu --set full --clock-control none --import-source on -k regex:rov6_step_kernel --launch-skip 8 -c 1 -o $O/rov6_sp python bench.py --no-cpu --no-extra --steps 5 --warmup 5 --graph 0 --action-mode setpoint > $O/ncu_sp.log 2>&1
python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
ls $O
