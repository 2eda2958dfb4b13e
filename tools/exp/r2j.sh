#!/bin/bash
# round 2, GPU call J: two-environments-per-thread auv_step kernel
O=gpurun_out/r2j; mkdir -p $O
P=$PWD/marinevehiclereinforcementlearning_b200
timeout 900 python -m pytest tests/test_auv_gpu.py tests/test_vec_tools_gpu.py tests/test_housekeeping_gpu.py tests/test_awkward_sizes_gpu.py tests/test_full_size_gpu.py -q -rA > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
grep -E "passed|failed|rc=|^FAILED|^E  " $O/pytest.log | tail -8
A="python bench.py --workload auv --steps 500 --warmup 50"
MVRL_LIB=$P/libmvrl_r1.so $A > $O/auv_r1.json 2>> $O/err.log
MVRL_AUV_NO_X2=1 $A > $O/auv_one.json 2>> $O/err.log
$A > $O/auv_x2.json 2>> $O/err.log
$A --envs 1048576 > $O/auv_x2_1m.json 2>> $O/err.log
MVRL_AUV_NO_X2=1 $A --envs 1048576 > $O/auv_one_1m.json 2>> $O/err.log
ncu --set full --clock-control none --import-source on -k regex:auv_step --launch-skip 280 -c 1 -o $O/auv python bench.py --workload auv --steps 20 --warmup 270 --graph 0 > $O/ncu_auv.log 2>&1
tail -3 $O/err.log
ls $O
