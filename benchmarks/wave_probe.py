import sys, os, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/benchmarks")
import numpy as np, torch
import kernex_b200 as kex
from kernex_b200 import _lib
from run_configs import timeit
dev = torch.device("cuda")
for shape in [(32, 8192), (64, 8192), (128, 8192), (4096, 256)]:
    x = torch.randn(shape, device=dev)
    for name, fn in [("gs4", lambda u: 0.25 * (u[0, -1] + u[-1, 0] + u[0, 1] + u[1, 0])),
                     ("new_only", lambda u: 0.5 * (u[0, -1] + u[-1, 0])),
                     ("row_only", lambda u: 0.5 * u[0, -1] + 0.25 * u[0, 0])]:
        f = kex.kscan(kernel_size=(3, 3), padding="same", relative=True)(fn)
        f(x)
        med, _ = timeit(lambda: f(x), 3, warm=1, inner=1)
        steps = shape[1] + shape[0] - 1 if name != "row_only" else shape[1]
        print(shape, name, _lib.last_kernel(), f"{med:.3f} ms", f"{med*1e3/steps:.3f} us/step", flush=True)
x = torch.randn((200000,), device=dev)
f = kex.kscan(kernel_size=(3,), padding="same", relative=True)(lambda u: 0.5 * u[-1] + u[0])
f(x); med, _ = timeit(lambda: f(x), 3, warm=1, inner=1)
print("1-D 200000", _lib.last_kernel(), f"{med:.3f} ms {med*1e3/200000:.3f} us/step")
