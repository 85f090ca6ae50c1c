#!/usr/bin/env python
"""Host-side cost of one conv-as-kmap call with device weights on a small image (the reference's benchmark harness is
launch-latency bound on every shape)."""
import cProfile, pstats, io, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kernex_b200 as kex

C, H = 8, 256
x = torch.randn((C, H, H), device="cuda")
w = torch.randn((C, 3, 3), device="cuda")
f = kex.kmap(kernel_size=(C, 3, 3), padding=("valid", "same", "same"))(lambda v, w: np.sum(v * w))
for _ in range(50):
    f(x, w)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000):
    f(x, w)
torch.cuda.synchronize()
print("us per call:", (time.perf_counter() - t0) / 2000 * 1e6)
pr = cProfile.Profile()
pr.enable()
for _ in range(2000):
    f(x, w)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue()[:4500])
