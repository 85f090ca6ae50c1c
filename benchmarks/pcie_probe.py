#!/usr/bin/env python
"""PCIe ceiling for the e2e arm: pinned H2D, D2H, and both directions at once (1 GiB each)."""
import time, torch
n = 1 << 28
h_in = torch.empty(n, dtype=torch.float32, pin_memory=True).normal_()
h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
d_in = torch.empty(n, dtype=torch.float32, device="cuda")
d_out = torch.randn(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
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
for name, a, b in [("H2D", 1, 0), ("D2H", 0, 1), ("both", 1, 1)]:
    run(a, b, 2)
    dt = run(a, b)
    print(f"{name}: {dt*1e3:.2f} ms per GiB direction -> {n*4/dt/1e9:.1f} GB/s per direction")
# chunked 32 MiB slabs both ways (what hostpipe does)
chunk = 8 << 20
def run_chunked(reps=3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for o in range(0, n, chunk):
            with torch.cuda.stream(s1):
                d_in[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[o:o + chunk].copy_(d_out[o:o + chunk], non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
run_chunked(1)
dt = run_chunked()
print(f"both, 32 MiB chunks: {dt*1e3:.2f} ms -> {n*4/dt/1e9:.1f} GB/s per direction")
