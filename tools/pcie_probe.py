"""Diagnostic: pinned-memory PCIe copy rates on the bench box (upper bound for bench.py's e2e leg).
    python tools/pcie_probe.py"""
import time

import torch

dev = torch.device("cuda:0")
n = 1 << 20
h_in = torch.empty(n * 8, dtype=torch.float32).pin_memory()
h_out = torch.empty(n * 10 + n // 4, dtype=torch.float32).pin_memory()
d_in, d_out = torch.empty_like(h_in, device=dev), torch.empty_like(h_out, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


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
t = timeit(h2d); print("H2D  %.1f MB  %.3f ms  %.1f GB/s" % (mb_in, t * 1e3, mb_in / t / 1e3))
t = timeit(d2h); print("D2H  %.1f MB  %.3f ms  %.1f GB/s" % (mb_out, t * 1e3, mb_out / t / 1e3))
t = timeit(both); print("both %.1f MB  %.3f ms  %.1f GB/s aggregate -> %.3g env-steps/s bound" % (mb_in + mb_out, t * 1e3, (mb_in + mb_out) / t / 1e3, n / t))
