#!/bin/bash
# round 2, GPU call P: the driver's sequence on the tree after the SPOD / replay additions and the auv variants' retirement
O=gpurun_out/r2p; mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
tail -3 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
python tools/sanitize_smoke.py > $O/sanitize_smoke_plain.log 2>&1; echo "awkward sizes rc=$?"
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python bench.py --workload auv --steps 500 --warmup 50 --envs 524288 > $O/auv_512k.json 2>> $O/err.log
python bench.py --workload auv --steps 500 --warmup 50 --envs 2097152 > $O/auv_2m.json 2>> $O/err.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv python bench.py --no-cpu --steps 5 --warmup 3 --extra-steps 5 --rollout-len 8 > $O/ncu_list.log 2>&1
ncu --set full --clock-control none -k regex:rov6_step_kernel --launch-skip 8 -c 1 -o $O/rov6_step python bench.py --no-cpu --no-extra --steps 5 --warmup 5 --graph 0 > $O/ncu_step.log 2>&1
ncu --set full --clock-control none -k regex:flow_reconstruct -c 1 -o $O/spod python tools/spod_time.py > $O/ncu_spod.log 2>&1
for f in $O/auv_*.json; do python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1].split('/')[-1], '%.4g'%d['value'], '%.2f us'%(d['ms_per_step']*1e3))
PY
done
ls $O
