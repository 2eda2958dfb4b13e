#!/bin/bash
# round 2, GPU call B: full GPU suite on the new kernels (pose-increment PID, mask selects, small-shard shape), the new
# bench line (extras + strong keys), small-shard comparison, ncu captures of the rpm / set-point / auv kernels
O=gpurun_out/r2b; mkdir -p $O
python -m pytest tests -m gpu -q -rA > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
grep -E "passed|failed|rc=" $O/pytest.log | tail -3
python bench.py --steps 300 --warmup 20 --cpu-steps 20 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
B="python bench.py --no-cpu --no-extra --steps 300 --warmup 20"
for e in 131072 65536 262144 524288; do
  $B --envs $e > $O/small_$e.json 2>> $O/err.log
  MVRL_NO_SMALL_SHAPE=1 $B --envs $e > $O/nosmall_$e.json 2>> $O/err.log
done
$B --envs 131072 --action-mode setpoint > $O/small_sp_131072.json 2>> $O/err.log
MVRL_NO_SMALL_SHAPE=1 $B --envs 131072 --action-mode setpoint > $O/nosmall_sp_131072.json 2>> $O/err.log
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:rov6_step_kernel --launch-skip 8 -c 1 -o $O/rov6_rpm python bench.py --no-cpu --no-extra --steps 5 --warmup 5 --graph 0 > $O/ncu_rpm.log 2>&1
$NCU -k regex:rov6_step_kernel --launch-skip 8 -c 1 -o $O/rov6_sp python bench.py --no-cpu --no-extra --steps 5 --warmup 5 --graph 0 --action-mode setpoint > $O/ncu_sp.log 2>&1
$NCU -k regex:auv_step_kernel --launch-skip 280 -c 1 -o $O/auv python bench.py --workload auv --steps 20 --warmup 270 --graph 0 > $O/ncu_auv.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --no-cpu --steps 5 --warmup 3 --extra-steps 5 > $O/ncu_list.log 2>&1
ls -la $O
