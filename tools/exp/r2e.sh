#!/bin/bash
# round 2, multi-GPU call: usage  bash tools/exp/r2e.sh "1 2" | "1 2 4 8"   (rank counts to run)
NS=${1:-"1 2"}
O=gpurun_out/r2e; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > $O/topo.txt 2>&1
for n in $NS; do
  $TR --nproc-per-node $n --master-port $((29600 + n)) tools/pcie_probe.py > $O/pcie_$n.json 2>> $O/err.log
done
for n in $NS; do
  if [ $n -eq 1 ]; then continue; fi
  $TR --nproc-per-node $n --master-port $((29700 + n)) bench.py --gpus $n --steps 300 --warmup 20 > $O/bench_$n.json 2>> $O/err.log; echo "bench $n rc=$?"
  $TR --nproc-per-node $n --master-port $((29800 + n)) bench.py --gpus $n --steps 300 --warmup 20 --scaling strong --no-extra > $O/bench_strong_$n.json 2>> $O/err.log; echo "strong $n rc=$?"
done
tail -5 $O/err.log
ls $O
