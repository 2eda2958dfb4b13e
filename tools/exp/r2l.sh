#!/bin/bash
# round 2, GPU call L: SPOD reconstruction kernel + replay buffer against the reference-generated goldens; legacy suite after the x2 removal
O=gpurun_out/r2l; mkdir -p $O
timeout 900 python -m pytest tests/test_spod_gpu.py tests/test_vec_tools_gpu.py tests/test_auv_gpu.py tests/test_housekeeping_gpu.py -q -rA > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
grep -E "passed|failed|rc=|^FAILED|^E  " $O/pytest.log | tail -12
python tools/spod_time.py > $O/spod_time.json 2> $O/spod_time.err; cat $O/spod_time.json; tail -3 $O/spod_time.err
