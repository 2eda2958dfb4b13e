#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_awkward_sizes_gpu.py tests/test_rov6_gpu.py -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/r1_chk.json 2> gpurun_out/r1_chk.err; echo rc=$? lines=$(wc -l < gpurun_out/r1_chk.json)
head -c 200 gpurun_out/r1_chk.json; echo
