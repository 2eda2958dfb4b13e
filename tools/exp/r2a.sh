#!/bin/bash
# round 2, GPU call A: GPU test suite, parity probe (pose-increment PID vs literal), launch-shape / unroll variants
O=gpurun_out/r2a; mkdir -p $O
P=$PWD/marinevehiclereinforcementlearning_b200
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/gpu.txt
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
python tools/exp/r2_parity_probe.py > $O/probe.log 2>&1
MVRL_LIB=$P/libmvrl_nodp.so python tools/exp/r2_parity_probe.py nodp > $O/probe_nodp.log 2>&1
B="python bench.py --no-cpu --steps 200 --warmup 20"
$B > $O/base_rpm_1m.json 2> $O/err.log
$B --envs 131072 > $O/base_rpm_128k.json 2>> $O/err.log
$B --envs 65536 > $O/base_rpm_64k.json 2>> $O/err.log
$B --action-mode setpoint > $O/base_sp_1m.json 2>> $O/err.log
$B --action-mode force > $O/base_force_1m.json 2>> $O/err.log
for e in 1048576 131072 65536; do MVRL_LIB=$P/libmvrl_b64.so $B --envs $e > $O/b64_rpm_$e.json 2>> $O/err.log; done
MVRL_LIB=$P/libmvrl_b64.so $B --action-mode setpoint > $O/b64_sp_1m.json 2>> $O/err.log
for v in u2 u1 u2m2 u4m2 nodp; do MVRL_LIB=$P/libmvrl_$v.so $B --action-mode setpoint > $O/${v}_sp_1m.json 2>> $O/err.log; done
for v in u2 u1; do MVRL_LIB=$P/libmvrl_$v.so $B > $O/${v}_rpm_1m.json 2>> $O/err.log; done
tail -3 $O/pytest.log
