#!/bin/bash
# round 2, GPU call AI (the last one of the round): the bare-MUFU reciprocal in the kinematics (rov6_model.cuh: rcp_mufu) -
# the driver's sequence on the changed tree, then a same-box A/B against the library built from the previous commit
# (libmvrl_head.so), the default bench lines, and the one-environment-per-thread kernel on the 8-GPU shard size
O=gpurun_out/r2ai; mkdir -p $O
H=marinevehiclereinforcementlearning_b200/libmvrl_head.so
timeout 420 python -m pytest tests -x -q -m gpu --durations=12 > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -4 $O/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
B="timeout 100 python bench.py --no-cpu --no-extra --steps 300 --warmup 20"
for r in 1 2; do
  MVRL_LIB=$H $B > $O/ab_head_rpm_$r.json 2>> $O/err.log
  $B > $O/ab_new_rpm_$r.json 2>> $O/err.log
done
for m in setpoint force; do
  MVRL_LIB=$H $B --action-mode $m > $O/ab_head_$m.json 2>> $O/err.log
  $B --action-mode $m > $O/ab_new_$m.json 2>> $O/err.log
done
for f in $O/ab_*.json; do python - $f <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], '%.4g env-steps/s' % d['value'], '%.2f us' % (d['ms_per_step'] * 1e3))
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
done
timeout 400 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
# one environment per thread (scalar FFMA, twice the warps) on small shards: never measured below 1 Mi environments
for n in 131072 65536; do
  MVRL_NO_X2=1 $B --envs $n > $O/nox2_$n.json 2>> $O/err.log
  $B --envs $n > $O/x2_$n.json 2>> $O/err.log
done
for f in $O/nox2_*.json $O/x2_*.json; do python - $f <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], '%.4g env-steps/s' % d['value'], '%.2f us' % (d['ms_per_step'] * 1e3))
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
done
timeout 200 ncu --set full --clock-control none -k regex:rov6_step_kernel --launch-skip 8 -c 1 -o $O/rov6_step python bench.py --no-cpu --no-extra --steps 5 --warmup 5 --graph 0 > $O/ncu_step.log 2>&1; echo "ncu rc=$?"
tail -3 $O/err.log
ls $O
