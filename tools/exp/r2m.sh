#!/bin/bash
# round 2, GPU call M: SPOD reconstruction on the fp64 tensor cores (DMMA)
O=gpurun_out/r2m; mkdir -p $O
timeout 600 python -m pytest tests/test_spod_gpu.py -q -rA > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
grep -E "passed|failed|rc=|^FAILED|^E  " $O/pytest.log | tail -12
python tools/spod_time.py > $O/spod_time.json 2> $O/spod_time.err; cat $O/spod_time.json; tail -3 $O/spod_time.err
