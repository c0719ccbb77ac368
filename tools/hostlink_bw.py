#!/usr/bin/env python3
"""Developer tool: pinned host <-> device copy bandwidth of this box (the bound of bench.py's `e2e`)."""
import time
import torch

n = 512 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


for _ in range(2):
    a = run(True, False)
    b = run(False, True)
    c = run(True, True)
    print("H2D alone %.1f GB/s | D2H alone %.1f GB/s | both: H2D %.1f + D2H %.1f GB/s (512 MiB in, 256 MiB out per rep)" % (
        n / a / 1e9, n / 2 / b / 1e9, n / c / 1e9, n / 2 / c / 1e9))
