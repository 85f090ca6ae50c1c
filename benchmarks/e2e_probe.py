#!/usr/bin/env python
"""e2e arm of bench.py alone (pinned host grid in, host result out), for tuning the slab pipeline."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kernex_b200 as kex
N = 16384
lap = kex.kmap(kernel_size=(3, 3), padding="valid", relative=True)(
    lambda x: x[1, 0] + x[0, -1] - 4 * x[0, 0] + x[0, 1] + x[-1, 0])
pinned = torch.empty((N, N), dtype=torch.float32, pin_memory=True)
x = pinned.numpy()
x[...] = np.random.default_rng(0).standard_normal((N, N), dtype=np.float32)
for _ in range(3):
    y = lap(x)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    y = lap(x)
dt = (time.perf_counter() - t0) / 10
print(f"slab {os.environ.get('KX_HOSTPIPE_SLAB_MB', '32')} MiB: {dt*1e3:.2f} ms/step  {(N-2)**2/dt/1e9:.2f} GLUP/s")
