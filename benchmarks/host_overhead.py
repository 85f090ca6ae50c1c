#!/usr/bin/env python
"""Host-side cost of one kmap call on a tiny device-resident input (BASELINE config 1 is launch-latency bound)."""
import cProfile, pstats, io, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kernex_b200 as kex

x = torch.arange(1, 1_000_001, dtype=torch.int32, device="cuda")
f = kex.kmap(kernel_size=(3,))(lambda v: np.sum(v))
for _ in range(50):
    f(x)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000):
    f(x)
torch.cuda.synchronize()
print("us per call:", (time.perf_counter() - t0) / 2000 * 1e6)
pr = cProfile.Profile()
pr.enable()
for _ in range(2000):
    f(x)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
print(s.getvalue()[:3500])
