"""Diagnostic: pinned-memory PCIe copy rates on the bench box - the upper bound of bench.py's e2e leg - for ONE GPU or
for N GPUs driven CONCURRENTLY (one process per GPU, like the bench):

    python tools/pcie_probe.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/pcie_probe.py

Every rank copies the bench's per-step host traffic (1 Mi environments: 33.6 MB of actions up, 38.8 MB of observations
+ done flags down) H2D only, D2H only and both at once on two streams, between barriers; rank 0 prints one JSON line
with the per-rank and the aggregate rates and the resulting bound on e2e env-steps/s.  If the aggregate duplex rate
stops growing with N, the GPUs share host-side bandwidth (PCIe switch uplinks / root complex / host memory)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
numa = None
try:
    import bench
    numa = bench.bind_to_gpu_numa_node(local)
except Exception:
    pass
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)

n = 1 << 20
h_in = torch.empty(n * 8, dtype=torch.float32).pin_memory()
h_out = torch.empty(n * 9 + n // 4, dtype=torch.float32).pin_memory()
d_in, d_out = torch.empty_like(h_in, device=dev), torch.empty_like(h_out, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / reps
    barrier()
    return t


def h2d():
    d_in.copy_(h_in, non_blocking=True)


def d2h():
    h_out.copy_(d_out, non_blocking=True)


def both():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


mb_in, mb_out = h_in.numel() * 4 / 1e6, h_out.numel() * 4 / 1e6
mine = torch.tensor([mb_in / timeit(h2d) / 1e3, mb_out / timeit(d2h) / 1e3, (mb_in + mb_out) / timeit(both) / 1e3,
                     float(-1 if numa is None else numa)], dtype=torch.float64, device=dev)
if world > 1:
    allr = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
else:
    allr = [mine]
if rank == 0:
    rows = [r.tolist() for r in allr]
    duplex = [r[2] for r in rows]
    out = {"gpus": world, "h2d_mb": mb_in, "d2h_mb": mb_out,
           "per_gpu_GBps": [{"h2d": round(r[0], 1), "d2h": round(r[1], 1), "duplex": round(r[2], 1), "numa_node": int(r[3])} for r in rows],
           "aggregate_duplex_GBps": round(sum(duplex), 1),
           # the slowest rank sets the step time of a synchronous job: bound = N envs-per-GPU / (bytes per GPU / its duplex rate)
           "e2e_bound_env_steps_per_s": world * n / ((mb_in + mb_out) / 1e3 / min(duplex))}
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
