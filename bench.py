#!/usr/bin/env python
"""Benchmark of the batched BlueROV2 6DoF env step (BASELINE.json metric:
"BlueROV2 6DoF env-steps/sec ... (1M envs); % of FP pipe peak").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config 3 of BASELINE.json): BlueROV2 Heavy 6DoF,
fp32, 1 048 576 environments PER GPU (weak scaling: the path shards by
environment with no data-path collective), direct thruster-rpm actions
~U(-3500, 3500) (seed 1234 + rank), dt = 0.2 s as nSub = 8 fixed RK4 sub-steps,
maxSteps = 250 with auto-reset.  One "step" = one launch of the fused step
kernel over the whole batch.  Prints ONE JSON line (rank 0).

`--impl reference` times the reference's CPU implementation of the same path:
the reference is pure Python (no compiled artefact can be built from it and
/root/reference does not exist on the GPU box), so this arm runs the C port of
it (oracle/mvrl_oracle.c, pinned to vectors produced by the unmodified
reference) on all host threads, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "BlueROV2 6DoF env-steps/sec"
UNIT = "env-steps/s"
ENVS_PER_GPU = 1 << 20
N_SUB = 8
DT = 0.2
MAX_STEPS = 250
# algorithmic work of one 6DoF env step, direct-rpm mode
FLOP_PER_ENV_STEP = 1600 * N_SUB + 60          # FMA = 2, other fp ops = 1, libm calls not counted
BYTES_PER_ENV_STEP_F32 = 177                   # state r/w, action r, obs/reward/done w, counter r/w
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12
NCU_TRAFFIC_BYTES = 176.3e6                    # measured DRAM bytes of one 1 Mi-env launch (profiles/r1_z_rov6_step_ncu_full_summary.txt)


ACTION_SCALE = {"rpm": 3500.0, "force": 40.0, "setpoint": 1.0}
ACTION_DIM = {"rpm": 8, "force": 6, "setpoint": 6}
# derivative = 360 flop (rpm) / 610 (PID set-point; force mode has the allocation but no PID: 538)
DERIV_FLOP = {"rpm": 360, "force": 538, "setpoint": 610}


def flop_per_env_step(mode, n_sub):
    return (4 * DERIV_FLOP[mode] + 160) * n_sub + 60


def bytes_per_env_step(mode, w):
    b = 12 * w * 2 + ACTION_DIM[mode] * w + 9 * w + w + 1 + 8          # state r/w, action r, obs/reward/done w, counter r/w
    if mode == "setpoint":
        b += 13 * w * 2 + 6 * w * 2 + 6 * w                            # PID state r/w, set-point r/w, path r
    return b


def workload_name(envs, mode="rpm", dtype="f32", n_sub=N_SUB):
    return ("rov6_step %s: BlueROV2 Heavy 6DoF, %d envs/GPU, %s actions U(-%g,%g), dt=%.1f as nSub=%d RK4, "
            "maxSteps=%d auto-reset" % (dtype, envs, mode, ACTION_SCALE[mode], ACTION_SCALE[mode], DT, n_sub, MAX_STEPS))


# --------------------------------------------------------------------------
# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on the first
# collective), so file descriptor 1 is pointed at stderr for the whole run and the line goes to the saved descriptor.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    text = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, text)
    else:
        os.write(_REAL_STDOUT, text)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                try:
                    sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no sample inside the timed region"], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


def bind_to_gpu_numa_node(index):
    """Pin this rank's host threads (and with them its first-touch / pinned allocations) to the CPUs of the NUMA node the
    GPU hangs off: the e2e leg moves 76 MB per step between host memory and the GPU, and with one rank per GPU the
    cross-socket hop otherwise becomes the limiter.  Returns the node, or None when the topology is not exposed."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------
def host_threads():
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which must not throttle the CPU arm)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_port_rate(n_envs, n_steps, threads=0, seed=1234):
    """env-steps/s of the C port of the reference's step loop on the host."""
    from oracle import c_oracle as c
    from oracle import oracle_np as o
    threads = threads or host_threads()
    env = c.Rov6EnvC(n_envs, mode=o.MODE_RPM, max_steps=MAX_STEPS, n_sub=N_SUB, dt=DT, auto_reset=True, seed=seed, threads=threads)
    env.reset()
    rng = np.random.default_rng(seed)
    acts = rng.uniform(-3500.0, 3500.0, (4, n_envs, 8))
    env.step(acts[0])  # warm-up (thread pool, page faults)
    t0 = time.perf_counter()
    for k in range(n_steps):
        env.step(acts[k % 4])
    dt = time.perf_counter() - t0
    return n_envs * n_steps / dt, dt, threads


def python_port_rate(seconds=3.0):
    """env-steps/s of the numpy restatement stepped ONE environment at a time -
    the shape of the reference's own loop (per-env Python/numpy, one core)."""
    from oracle import oracle_np as o
    env = o.Rov6EnvOracle(1, mode=o.MODE_RPM, max_steps=10 ** 9, n_sub=N_SUB, dt=DT)
    env.reset(initial_setpoint=np.zeros(6))
    rng = np.random.default_rng(0)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        env.step(rng.uniform(-3500, 3500, (1, 8)))
        n += 1
    return n / (time.perf_counter() - t0)


def run_reference(args, rank, world):
    if rank != 0:
        return
    n_envs = 32768
    from oracle import c_oracle as c
    from oracle import oracle_np as o
    threads = host_threads()
    env = c.Rov6EnvC(n_envs, mode=o.MODE_RPM, max_steps=MAX_STEPS, n_sub=N_SUB, dt=DT, auto_reset=True, seed=1234, threads=threads)
    env.reset()
    rng = np.random.default_rng(1234)
    acts = rng.uniform(-3500.0, 3500.0, (4, n_envs, 8))
    for k in range(args.warmup):
        env.step(acts[k % 4])
    t0 = time.perf_counter()
    for k in range(args.steps):
        env.step(acts[k % 4])
    t_total = time.perf_counter() - t0
    rates = [0] * args.steps
    value = n_envs * args.steps / t_total
    sample = "%d envs x 1 env step per bench step (nSub=%d RK4), all host threads" % (n_envs, N_SUB)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_total / len(rates), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(ENVS_PER_GPU), "l2": "n/a (CPU arm)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# --------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from marinevehiclereinforcementlearning_b200 import BlueROV2Heavy6DoFVecEnv, _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if args.bind_numa else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.envs
    mode, n_sub = args.action_mode, args.n_sub
    tdtype = torch.float64 if args.dtype == "f64" else torch.float32
    w = 8 if args.dtype == "f64" else 4
    na = ACTION_DIM[mode]
    env = BlueROV2Heavy6DoFVecEnv(n, action_mode=mode, dtype=tdtype, device=dev, dt=DT, maxSteps=args.max_steps, n_sub=n_sub,
                                  seed=1234, env_id0=rank * n, auto_reset=True, fast_math=bool(args.fast_math),
                                  record_terminal_obs=False, collect_stats=not args.no_stats)
    env.reset()
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    n_act = 4  # rotating action batches: 4 x 32 MiB on top of 125 MB touched per step > 126 MB L2
    acts = [(torch.rand((na, env.ld), generator=gen, device=dev, dtype=tdtype) * 2 - 1) * ACTION_SCALE[mode] for _ in range(n_act)]
    # stagger episode phase like a long-running job: env i starts at iStep = i % maxSteps
    env._istep.copy_((torch.arange(env.ld, device=dev) % MAX_STEPS).to(torch.int32))

    def one_step(k):
        env._bufs.action = acts[k % n_act].data_ptr()
        env.step_async()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(args.warmup):
        one_step(k)
    barrier()
    graph = None
    if args.graph:  # replay the K timed launches as one CUDA graph (no per-launch host cost)
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for k in range(args.steps):
                    one_step(k)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph.replay()   # warm-up replay
        barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    if graph is not None:
        graph.replay()
    else:
        for k in range(args.steps):
            one_step(k)
    e1.record()
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t0, t1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    value = world * n * args.steps / (ms * 1e-3)
    ms_per_step = ms / args.steps
    stats = None if args.no_stats else env.episode_stats(reset=True)  # K5; all-reduced over NCCL when world > 1 (off the timed path)

    # ---- end to end through the public API with HOST buffers -----------------
    # BlueROV2Heavy6DoFVecEnv.step_host -> mvrl_rov6_step_host: pinned host [N, 8] actions in, pinned host
    # obs [N, 9] / reward [N] / done [N] out, every step; upload / step / download pipelined over chunks.
    h_act = [a[:, :n].T.contiguous().cpu().pin_memory() for a in acts[:2]]          # [N, A] like a VecEnv caller
    env._bufs.action = env._action.data_ptr()

    def e2e_step(k):
        env.step_host(h_act[k % 2], chunks=args.e2e_chunks)                # returns when the host tensors are complete

    e2e_steps = max(3, min(args.steps, 50))
    for k in range(3):
        e2e_step(k)
    barrier()
    w0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(k)
    barrier()
    e2e_s = time.perf_counter() - w0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / float(t[0])
    h2d = n * na * w
    d2h = n * (9 * w + w + 1)
    # launches inside the device-timed region: one fused step kernel per bench step
    e2e_pieces = _lib.load().mvrl_host_chunk_count(n, args.e2e_chunks)
    e2e_launches_per_step = 3 * e2e_pieces

    if rank == 0:
        peaks, peak_src = measured_peaks()
        fp32_peak = _lib.measure_fma_peak(_lib.F32, local_rank)   # K6, measured now on this GPU
        fp64_peak = _lib.measure_fma_peak(_lib.F64, local_rank, iters=1024)
        per_gpu_rate = n / (ms_per_step * 1e-3)
        flop, nbytes = flop_per_env_step(mode, n_sub), bytes_per_env_step(mode, w)
        ach_tflops = per_gpu_rate * flop / 1e12
        ach_gbs = per_gpu_rate * nbytes / 1e9
        fp_peak = fp64_peak if args.dtype == "f64" else fp32_peak
        roofline = {"bound": "fp64" if args.dtype == "f64" else "fp32", "achieved": ach_tflops, "peak": fp_peak, "unit": "TFLOP/s",
                    "frac": ach_tflops / fp_peak, "traffic": NCU_TRAFFIC_BYTES if (mode, args.dtype, n_sub, n) == ("rpm", "f32", N_SUB, ENVS_PER_GPU) else None,
                    "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one rov6_step launch, ncu --set full, profiles/r1_z_rov6_step_ncu_full_summary.txt",
                    "peak_source": "FMA-chain microbenchmark (mvrl_measure_fma_peak: operands from uniform registers) run on this GPU in this "
                    "process; nominal fp32 %.1f. With three distinct register operands FFMA sustains only 0.61 inst/clk/SMSP "
                    "(45.7 TFLOP/s, tools/ffma_regs.cu)" % NOMINAL_FP32_TFLOPS,
                    "fp32_peak_tflops": fp32_peak, "fp64_peak_tflops": fp64_peak, "flop_per_env_step": flop,
                    # what the shipped kernel actually executes (packed FFMA2 = 2 FMA; counted from its SASS with
                    # tools/sass_operands.py: 365 FFMA2 + 158 FMUL2 + 37 FADD2 + 16 scalar per sub-step and pair of environments):
                    # folding / hoisting / anchored trig make it fewer flops than the frozen algorithmic count, hence frac ~ 1
                    "executed_flop_per_env_step": (933 * n_sub + 100) if (mode, args.dtype) == ("rpm", "f32") else None,
                    "executed_frac": (per_gpu_rate * (933 * n_sub + 100) / 1e12 / fp_peak) if (mode, args.dtype) == ("rpm", "f32") else None,
                    "hbm": {"achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach_gbs / peaks["hbm_gbs"],
                            "bytes_per_env_step": nbytes, "peak_source": peak_src}}
        cpu = None
        if not args.no_cpu and world == 1:   # the CPU baseline is reported at N = 1 only
            rate, secs, used = cpu_port_rate(32768, args.cpu_steps)
            cpu = {"value": rate, "unit": UNIT, "cores": used, "kind": "port",
                   "sample": "%d envs x %d env steps of the same workload, C port of the reference loop, %d threads, %.1f s" %
                             (32768, args.cpu_steps, used, secs),
                   "python_port_one_core": python_port_rate(2.0)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": workload_name(n, mode, args.dtype, n_sub), "envs_per_gpu": n, "envs_total": n * world, "n_sub": n_sub,
                           "action_mode": mode, "fast_math": bool(args.fast_math), "cuda_graph": bool(args.graph),
                           "two_envs_per_thread_ffma2": os.environ.get("MVRL_NO_X2", "0") != "1" and args.dtype == "f32", "parallelism": "env-sharded x%d, no collective on the step path" % world,
                           "l2": "inputs larger than L2: ~125 MB touched per step + 4 rotating 32 MiB action batches"},
                "roofline": roofline, "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                        "chunks": e2e_pieces, "host_numa_node": numa_node, "gpu_launches_per_step": e2e_launches_per_step,
                        "path": "BlueROV2Heavy6DoFVecEnv.step_host (mvrl_rov6_step_host): pinned host [N,8] actions -> pinned host obs/reward/done, "
                                "chunked H2D / transpose / fused step / transpose / D2H pipeline (obs by copy engine, reward + done stored into the pinned host arrays by the transpose kernel)"},
                "gpu_launches": args.steps, "clocks": clocks, "episode_stats": stats}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------
# secondary workloads (BASELINE configs 4 and 5); same JSON shape, selected with --workload
def _timed(dev, world, fn, steps, warmup):
    """W untimed + K timed calls of fn(k) between barriers; returns (ms total max over ranks, clocks)."""
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for k in range(warmup):
        fn(k)
    barrier()
    sampler = ClockSampler(dev.index)
    sampler.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    for k in range(steps):
        fn(k)
    e1.record()
    barrier()
    t1 = time.time()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), sampler.stop(t0, t1)


def run_auv(args, rank, local_rank, world):
    """Config 4: legacy AuvEnv, 262 144 envs per GPU, fp32, synthetic turbulence field [2000, 41, 61] scaled like
    verySimpleAuv.py:104, a ~ U(-1, 1)^3, noiseMag* = 0.1.  One step = one auv_step launch over the batch."""
    import torch
    import torch.distributed as dist
    from marinevehiclereinforcementlearning_b200 import AuvVecEnv
    from marinevehiclereinforcementlearning_b200.tag_00_Dec2023_simpleControlTurbulence import flowGenerator
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.envs if args.envs != ENVS_PER_GPU else 262144
    ltm = np.load(os.path.join(ROOT, "tests", "golden", "golden_legacy.npz"))["ltm"]   # the reference's ltm.npy (41 x 61 x 3)
    flow = flowGenerator.ReconstructedFlow.synthetic(lt_mean=ltm, nt=2000, seed=7, sigma=0.05, kind=args.field, dtype=torch.float32, device=dev)
    flow.scale(11., 1.0, 2.0, translate=(-1.65, -1.1))
    env = AuvVecEnv(n, flow, seed=1234, env_id0=rank * n, noiseMagCoeffs=0.1, noiseMagActuation=0.1, auto_reset=True,
                    dtype=torch.float32, record_terminal_obs=False)
    env.reset()
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    acts = [torch.rand((3, env.ld), generator=gen, device=dev) * 2 - 1 for _ in range(8)]

    def one_step(k):
        env._bufs.action = acts[k % 8].data_ptr()
        env.step_async()
    graph = None
    for k in range(3):
        one_step(k)
    if args.graph:
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for k in range(args.steps):
                    one_step(k)
        torch.cuda.current_stream(dev).wait_stream(side)
        ms, clocks = _timed(dev, world, lambda k: graph.replay() if k == 0 else None, 1, 0)
        ms, clocks = _timed(dev, world, lambda k: graph.replay(), 1, 1)
    else:
        ms, clocks = _timed(dev, world, one_step, args.steps, args.warmup)
    stats = env.episode_stats()
    if rank == 0:
        peaks, peak_src = measured_peaks()
        rate = n / (ms / args.steps * 1e-3)
        # state 6 r/w, action 3 r, obs 11 w, reward w, mults 11 r, target 2 r, err_o 3 r/w, ring 30 r + 3 w, return r/w, done 1, istep 8
        nbytes = (12 + 3 + 11 + 1 + 11 + 2 + 6 + 33 + 2) * 4 + 9
        line = {"metric": "legacy AuvEnv env-steps/sec", "value": world * rate, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "auv_step fp32: legacy verySimpleAuv, %d envs/GPU, field [2000,41,61,2] (%s), a~U(-1,1)^3" % (n, args.field),
                           "cuda_graph": bool(args.graph), "smem_staged_gather": os.environ.get("MVRL_AUV_NO_STAGE", "0") != "1",
                           "l2": "40 MB field is L2-resident by design; per-env arrays (87 MB / step) rotate through 8 action batches"},
                "roofline": {"bound": "hbm", "achieved": rate * nbytes / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": rate * nbytes / 1e9 / peaks["hbm_gbs"], "traffic": None, "bytes_per_env_step": nbytes,
                             "gathered_bytes_per_env_step_from_l2": 64, "peak_source": peak_src},
                "gpu_launches": args.steps, "clocks": clocks, "episode_stats": stats}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_rov3(args, rank, local_rank, world):
    """Config 1's model at scale: BlueROV2 Heavy 3DoF env, fp32, 1 Mi envs per GPU, dt = 0.2 as nSub RK4 sub-steps,
    auto-reset; --action-mode setpoint = the reference's Gym semantics (built-in PID), rpm = 4 thruster rpm."""
    import torch
    import torch.distributed as dist
    from marinevehiclereinforcementlearning_b200 import BlueROV2Heavy3DoFVecEnv
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, mode = args.envs, ("rpm" if args.action_mode == "rpm" else "setpoint")
    na, scale = (4, 3500.0) if mode == "rpm" else (3, 1.0)
    env = BlueROV2Heavy3DoFVecEnv(n, action_mode=mode, dtype=torch.float32, device=dev, dt=DT, maxSteps=MAX_STEPS, n_sub=args.n_sub,
                                  seed=1234, env_id0=rank * n, auto_reset=True, record_terminal_obs=False)
    env.reset()
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    acts = [(torch.rand((na, env.ld), generator=gen, device=dev) * 2 - 1) * scale for _ in range(8)]
    env._istep.copy_((torch.arange(env.ld, device=dev) % MAX_STEPS).to(torch.int32))

    def one_step(k):
        env._bufs.action = acts[k % 8].data_ptr()
        env.step_async()
    ms, clocks = _timed(dev, world, one_step, args.steps, args.warmup)
    stats = env.episode_stats()
    if rank == 0:
        rate = n / (ms / args.steps * 1e-3)
        line = {"metric": "BlueROV2 3DoF env-steps/sec", "value": world * rate, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "rov3_step fp32: BlueROV2 Heavy 3DoF, %d envs/GPU, %s actions, dt=%.1f as nSub=%d RK4, maxSteps=%d auto-reset"
                                       % (n, mode, DT, args.n_sub, MAX_STEPS)},
                "gpu_launches": args.steps, "clocks": clocks, "episode_stats": stats}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_rollout(args, rank, local_rank, world):
    """Config 5: rollout collection over the 6DoF env in the reference's Gym semantics (PID set-point actions):
    131 072 envs per GPU, policy MLP 9-128-128-128-6 (GELU, arch of legacy/main_00_sbl.py:100-105) + Gaussian head in
    PyTorch on the feature-major observation buffer, 128-step rollouts replayed as one CUDA graph, episode statistics
    all-reduced once per rollout.  Reported: env-steps/s including the policy."""
    import torch
    import torch.distributed as dist
    from marinevehiclereinforcementlearning_b200 import BlueROV2Heavy6DoFVecEnv
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = True
    n = args.envs if args.envs != ENVS_PER_GPU else 131072
    T = args.rollout_len
    env = BlueROV2Heavy6DoFVecEnv(n, action_mode="setpoint", dtype=torch.float32, device=dev, n_sub=args.n_sub, seed=1234,
                                  env_id0=rank * n, auto_reset=True, record_terminal_obs=False)
    env.reset()
    torch.manual_seed(1234 + rank)
    dims = [9, 128, 128, 128, 6]
    Ws = [torch.randn(dims[i + 1], dims[i], device=dev) / dims[i] ** 0.5 for i in range(4)]
    bs = [torch.zeros(dims[i + 1], 1, device=dev) for i in range(4)]
    log_std = torch.full((6, 1), -0.5, device=dev)
    ld = env.ld
    buf_obs = torch.empty((T, 9, ld), device=dev)
    buf_act = torch.empty((T, 6, ld), device=dev)
    buf_logp = torch.empty((T, ld), device=dev)
    buf_rew = torch.empty((T, ld), device=dev)
    buf_done = torch.empty((T, ld), dtype=torch.uint8, device=dev)

    Wt = [w.T.contiguous() for w in Ws]         # [in, out]
    b1 = [b.reshape(-1).contiguous() for b in bs]
    obs_nk = env._obs.T                          # [ld, 9] strided view of the SoA observation buffer: cuBLAS reads it in place

    def policy(x_nk):         # batch-major activations [N, k]; bias + GELU fused into the GEMM epilogue (cuBLASLt)
        h = x_nk
        for i in range(3):
            h = torch._addmm_activation(b1[i], h, Wt[i], use_gelu=True)
        return torch.tanh(torch.addmm(b1[3], h, Wt[3]))      # [N, 6]

    std = log_std.exp().reshape(1, 6)
    logp_const = float(-log_std.sum())

    def rollout():
        for t in range(T):
            buf_obs[t].copy_(env._obs)
            mean = policy(obs_nk)
            eps = torch.randn_like(mean)
            act = torch.addcmul(mean, eps, std).clamp_(-1., 1.)
            buf_logp[t].copy_((eps * eps).sum(1).mul_(-0.5).add_(logp_const))
            buf_act[t].copy_(act.T)             # back to the env's feature-major action layout
            env._bufs.action = buf_act[t].data_ptr()
            env.step_async()
            buf_rew[t].copy_(env._reward)
            buf_done[t].copy_(env._done)
    rollout()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            rollout()
    torch.cuda.current_stream(dev).wait_stream(side)
    stats_holder = {}

    def one_rollout(k):
        graph.replay()
        stats_holder["s"] = env.episode_stats()     # K5: device accumulators -> all-reduce (NCCL) -> host, once per rollout
    rollouts = max(2, args.steps // T)
    ms, clocks = _timed(dev, world, one_rollout, rollouts, max(1, args.warmup // T))
    if rank == 0:
        rate = n * T * rollouts / (ms * 1e-3)
        flop_env = flop_per_env_step("setpoint", args.n_sub)
        flop_policy = 2 * sum(dims[i] * dims[i + 1] for i in range(4))
        line = {"metric": "6DoF rollout collection env-steps/sec (policy included)", "value": world * rate, "unit": UNIT, "n_gpus": world,
                "steps": rollouts * T, "warmup": max(1, args.warmup // T) * T, "ms_per_step": ms / (rollouts * T), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 (policy matmuls TF32)", "data": "synthetic",
                "config": {"workload": "rollout: 6DoF set-point env + MLP 9-128-128-128-6 GELU Gaussian policy, %d envs/GPU, %d-step rollouts, "
                                       "nSub=%d, CUDA-graph replay, stats all-reduce per rollout" % (n, T, args.n_sub)},
                "flop_per_env_step": {"env": flop_env, "policy": flop_policy},
                "gpu_launches": rollouts * T, "clocks": clocks, "episode_stats": stats_holder.get("s")}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="environments per GPU")
    ap.add_argument("--fast-math", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-steps", type=int, default=100)
    ap.add_argument("--e2e-chunks", type=int, default=0, help="pieces of the host-buffer pipeline; 0 = the library's default")
    ap.add_argument("--action-mode", default="rpm", choices=["rpm", "force", "setpoint"], help="default rpm = BASELINE config 3")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--n-sub", type=int, default=N_SUB)
    ap.add_argument("--workload", default="rov6", choices=["rov6", "rov3", "auv", "rollout"],
                    help="rov6 = BASELINE config 3 (the metric); rov3 = the 3DoF env at scale; auv = config 4; rollout = config 5")
    ap.add_argument("--field", default="modes", choices=["modes", "noise"], help="auv: synthetic turbulence stand-in")
    ap.add_argument("--rollout-len", type=int, default=128)
    ap.add_argument("--max-steps", type=int, default=MAX_STEPS, help="episode length (diagnostics; default = the reference's 250)")
    ap.add_argument("--graph", type=int, default=1, help="1: the K timed launches are replayed as one CUDA graph; 0: K separate launches")
    ap.add_argument("--bind-numa", type=int, default=1, help="1: bind each rank to the CPUs of its GPU's NUMA node (host buffers of the e2e leg)")
    ap.add_argument("--no-stats", action="store_true", help="diagnostics: do not accumulate episode statistics in the step kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    capture_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == "rov3":
        run_rov3(args, rank, local_rank, world)
    elif args.workload == "auv":
        run_auv(args, rank, local_rank, world)
    elif args.workload == "rollout":
        run_rollout(args, rank, local_rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
